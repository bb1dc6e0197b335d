// Host logic of the C++ mirror's VoxelCachedLazyPRM::createRoadmap and TendonRobot::random_state, run WITHOUT
// a GPU: linked against tests/cpp/abi_standin_over_oracle.cpp instead of libirt_b200.so (see that file).
// Built and run by tests/test_abi_and_host.py.  The GPU run of the same checks is tests/cpp/test_host_mirror.cpp.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <set>
#include <vector>

#include "../../interactive-rate-tendons_b200/host/irt_host.hpp"
#include "../../oracle/tendon_oracle.h"

static int failures = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);           \
      failures++;                                                           \
    }                                                                       \
  } while (0)

#include "create_roadmap_checks.hpp"

int main() {
  tendon::TendonRobot robot;  // Robot B of the GPU test: 6 helical tendons, rotation + retraction
  robot.specs.dL = 0.003;
  robot.enable_retraction = true;
  robot.enable_rotation = true;
  for (int k = 0; k < 6; k++) {
    tendon::TendonSpecs t;
    t.C = {k * M_PI / 3, (k % 2 ? -1.0 : 1.0) * 2 * M_PI / robot.specs.L};
    t.D = {0.01};
    robot.tendons.push_back(t);
  }
  irt_robot_desc d = robot.desc();
  orc_robot orb;
  std::memcpy(&orb, &d, sizeof(orb));

  collision::VoxelOctree env_vox(128);
  env_vox.set_xlim(-0.21, 0.21); env_vox.set_ylim(-0.21, 0.21); env_vox.set_zlim(-0.21, 0.21);
  orc_grid og;
  std::memset(&og, 0, sizeof(og));
  og.Ng = 128;
  for (int a = 0; a < 3; a++) { og.lim[2 * a] = -0.21; og.lim[2 * a + 1] = 0.21; }
  og.inv_rot[0] = og.inv_rot[4] = og.inv_rot[8] = 1;
  orc_octree *oenv = orc_octree_new(&og);
  double c[3] = {0.05, 0.02, 0.12};
  orc_octree_add_sphere(oenv, c, 0.03);
  std::vector<uint8_t> xyz(3 * 4096);
  std::vector<uint64_t> bits(4096);
  int64_t nb = orc_octree_export(oenv, 4096, xyz.data(), bits.data());
  for (int64_t i = 0; i < nb; i++) env_vox.set_block(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], bits[i]);

  motion_planning::VoxelEnvironment venv;
  orc_space osp{0.02, 0.01, 0.0001};
  check_create_roadmap(robot, venv, env_vox, orb, og, oenv, osp);
  orc_octree_free(oenv);
  std::printf(failures ? "FAILED (%d)\n" : "createRoadmap host logic ok\n", failures);
  return failures ? 1 : 0;
}
