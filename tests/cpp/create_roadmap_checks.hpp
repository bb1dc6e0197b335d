// createRoadmap(N, opt) and TendonRobot::random_state of the C++ host mirror, checked against the CPU oracle.
// Part of tests/cpp/test_host_mirror.cpp, which runs against the real libirt_b200.so on the GPU box and, for the
// host logic, against tests/cpp/abi_standin_over_oracle.cpp where there is no GPU.
// Needs CHECK(), the mirror header and oracle/tendon_oracle.h from the including file.
#pragma once

// (VoxelCachedLazyPRM.cpp:1380-1561, tendon/TendonRobot.cpp:219-247)
static void check_create_roadmap(const tendon::TendonRobot &robot, const motion_planning::VoxelEnvironment &venv,
                               const collision::VoxelOctree &env_vox, const orc_robot &orb, const orc_grid &og,
                               const orc_octree *oenv, const orc_space &osp) {
  for (int i = 0; i < 50; i++) {
    auto s = robot.random_state();
    CHECK(s.size() == 8);
    for (int k = 0; k < 6; k++) CHECK(s[k] >= 0.0 && s[k] <= 20.0);
    CHECK(s[6] >= -M_PI && s[6] <= M_PI && s[7] >= 0.0 && s[7] <= robot.specs.L);
  }
  using PRM = motion_planning::VoxelCachedLazyPRM;
  PRM rm(robot, venv, env_vox);
  rm.setSeed(7);
  rm.createRoadmap(120, PRM::VoxelizeVertices | PRM::ValidateVertices | PRM::VoxelizeEdges | PRM::ValidateEdges);
  CHECK(rm.states().size() == 120 && rm.vertexValidity().size() == 120);
  CHECK(rm.edgeValidity().size() == rm.edges().size() && rm.edgeFlags().size() == rm.edges().size());
  // every accepted vertex: valid shape, misses the environment (the rejection rule of .cpp:1415-1443)
  for (size_t i = 0; i < rm.states().size(); i++) {
    std::vector<double> t(512), p(512 * 3);
    orc_fk_out fo;
    int n = orc_shape(&orb, rm.states()[i].data(), 512, t.data(), p.data(), nullptr, &fo);
    CHECK(orc_validity_flags(&orb, rm.states()[i].data(), &fo, p.data()) == 0);
    orc_octree *ov = orc_octree_new(&og);
    orc_voxelize_shape(&og, p.data(), n, ov);
    CHECK(orc_octree_collides(oenv, ov) != 1);
    orc_octree_free(ov);
    CHECK(rm.vertexValidity()[i] == PRM::VALIDITY_TRUE && rm.vertexFlags()[i] == 0);
    for (int c = 0; c < 3; c++) CHECK(std::fabs(rm.tipPositions()[3 * i + c] - p[3 * (n - 1) + c]) < 1e-9 * robot.specs.L);
  }
  // the edges: replay the connection loop (k nearest incl. the vertex itself, range-bounded, new vertices
  // in index order, no duplicates) and ask the oracle about every candidate
  std::set<std::pair<size_t, size_t>> have;
  std::vector<std::pair<size_t, size_t>> want;
  size_t n_candidates = 0;
  for (size_t v = 0; v < 120; v++) {
    std::vector<std::pair<double, size_t>> d;
    for (size_t i = 0; i < 120; i++) d.push_back({rm.distance(rm.states()[v], rm.states()[i]), i});
    std::sort(d.begin(), d.end());
    CHECK(d[0].second == v && d[0].first == 0.0);
    for (size_t j = 0; j < 5 && d[j].first <= rm.getRange(); j++) {
      const size_t nb = d[j].second;
      if (nb == v || !have.insert(std::minmax(v, nb)).second) continue;
      n_candidates++;
      orc_octree *ov = orc_octree_new(&og);
      orc_edge_out info;
      orc_voxelize_edge(&orb, &og, &osp, rm.states()[v].data(), rm.states()[nb].data(), nullptr, ov, &info);
      if (info.is_fully_valid && orc_octree_collides(oenv, ov) != 1) want.emplace_back(v, nb);
      orc_octree_free(ov);
    }
  }
  CHECK(want == rm.edges());
  for (auto v : rm.edgeValidity()) CHECK(v == PRM::VALIDITY_TRUE);
  for (auto f : rm.edgeFlags()) CHECK(f == 0);
  std::printf("createRoadmap: 120 vertices, %zu of %zu candidate edges kept\n", rm.edges().size(), n_candidates);
  CHECK(std::fabs(rm.maximumExtent() - std::sqrt(6 * 400.0) * (1.0 + 0.25 + 2.0)) < 1e-12);
  // growing: options pertain to what is added; N <= size is a no-op
  auto states0 = rm.states();
  auto edges0 = rm.edges();
  rm.createRoadmap(100);
  CHECK(rm.states().size() == 120 && rm.edges() == edges0);
  rm.createRoadmap(150, PRM::LazyRoadmap);
  CHECK(rm.states().size() == 150 && rm.edges().size() >= edges0.size());
  CHECK(std::equal(states0.begin(), states0.end(), rm.states().begin()));
  CHECK(std::equal(edges0.begin(), edges0.end(), rm.edges().begin()));
  for (size_t i = 0; i < 150; i++) CHECK(rm.vertexValidity()[i] == (i < 120 ? PRM::VALIDITY_TRUE : PRM::VALIDITY_UNKNOWN));
  for (size_t i = 0; i < rm.edges().size(); i++)
    CHECK(rm.edgeValidity()[i] == (i < edges0.size() ? PRM::VALIDITY_TRUE : PRM::VALIDITY_UNKNOWN));
  // clear*VoxelCache (.cpp:1814-1829): the sets go, the validity words stay; the next sweep rebuilds the caches
  rm.clearVoxelCache();
  CHECK(rm.vertexVoxelCacheSize() == 0 && rm.edgeVoxelCacheSize() == 0 && rm.vertexFlags().empty());
  CHECK(rm.vertexValidity()[0] == PRM::VALIDITY_TRUE && rm.edgeValidity()[0] == PRM::VALIDITY_TRUE);
  rm.precomputeValidity();
  CHECK(rm.vertexVoxelCacheSize() == 150 && rm.edgeVoxelCacheSize() == rm.edges().size());
  for (size_t i = 0; i < 120; i++) CHECK(rm.vertexValidity()[i] == PRM::VALIDITY_TRUE);
  for (size_t i = 0; i < edges0.size(); i++) CHECK(rm.edgeValidity()[i] == PRM::VALIDITY_TRUE);
  // a lazy roadmap from scratch touches no kernel-side state: no caches, nothing validated
  PRM lazy(robot, venv, env_vox);
  lazy.createRoadmap(40);
  CHECK(lazy.states().size() == 40 && lazy.vertexFlags().empty() && lazy.edgeFlags().empty());
  lazy.precomputeValidity();
  CHECK(lazy.vertexValidity().size() == 40 && lazy.edgeValidity().size() == lazy.edges().size());
  }
