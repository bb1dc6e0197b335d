// env_prims.h -- per-leaf-block arithmetic of Environment::voxelize's primitives, shared by
// the device kernel (env_prep.cu) and a host test harness (tests/cpp/test_env_prims_host.cpp) so that the
// cell-centre arithmetic can be checked against the oracle without a GPU.
//
// Reference: VoxelOctree::add_sphere / add_capsule (collision/VoxelOctree.cpp:434-515) mark every voxel whose
// CENTRE lies inside the object -- collides(Sphere, Point): |c - p|^2 <= r^2, collides(Capsule, Point): the
// same around interpolate(a, b, closest_t_segment(a, b, p)) (collision/collision.hxx:62-84,
// collision_primitives.h:17-49) -- over the blocks of the object's bounding box, after add_point() of the
// centre / end points (VoxelOctree.cpp:319-323).  A voxel centre that passes the test always lies in a block
// of that bounding box (centres sit half a cell away from block faces, the box is rounded by ulps), so testing
// a block against every object and OR-ing gives the same grid.
//
// Every floating-point operation is spelled out in the reference's order; on the device they are the
// round-to-nearest intrinsics (never contracted into FMAs), a host build needs -ffp-contract=off.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define EP_ADD(a, b) __dadd_rn((a), (b))
#define EP_SUB(a, b) __dsub_rn((a), (b))
#define EP_MUL(a, b) __dmul_rn((a), (b))
#define EP_DIV(a, b) __ddiv_rn((a), (b))
#else
#define EP_ADD(a, b) ((a) + (b))
#define EP_SUB(a, b) ((a) - (b))
#define EP_MUL(a, b) ((a) * (b))
#define EP_DIV(a, b) ((a) / (b))
#endif
#ifdef __CUDACC__
#define EP_HD __host__ __device__ __forceinline__
#else
#define EP_HD inline
#endif

// one object = 8 doubles: a[3], b[3], r, kind (0 = sphere with centre a, 1 = capsule a-b)
#define EP_PRIM_DOUBLES 8

EP_HD double ep_dot(double ax, double ay, double az, double bx, double by, double bz) {
  return EP_ADD(EP_ADD(EP_MUL(ax, bx), EP_MUL(ay, by)), EP_MUL(az, bz));  // Eigen's 3-vector dot order
}

// bits of leaf block (bx, by, bz) whose voxel centres lie inside the object; bit = i*16 + j*4 + k
// (VoxelOctree::bitmask, VoxelOctree.cpp:1501-1503).  lo = grid minimum, d = cell size.
EP_HD uint64_t ep_block_bits(const double *lo, const double *d, int bx, int by, int bz, const double *prim) {
  const double ax = prim[0], ay = prim[1], az = prim[2], r = prim[6];
  const bool capsule = prim[7] != 0.0;
  const double ex = capsule ? prim[3] : ax, ey = capsule ? prim[4] : ay, ez = capsule ? prim[5] : az;
  double cx[4], cy[4], cz[4];
  for (int i = 0; i < 4; i++) {  // xmin + dx * (ix + 0.5)
    cx[i] = EP_ADD(lo[0], EP_MUL(d[0], (double)((bx << 2) + i) + 0.5));
    cy[i] = EP_ADD(lo[1], EP_MUL(d[1], (double)((by << 2) + i) + 0.5));
    cz[i] = EP_ADD(lo[2], EP_MUL(d[2], (double)((bz << 2) + i) + 0.5));
  }
  // conservative reject against the object's bounding box, widened far beyond any rounding of the test below
  {
    const double mx = 1e-6 * d[0] + 1e-9 * r, my = 1e-6 * d[1] + 1e-9 * r, mz = 1e-6 * d[2] + 1e-9 * r;
    const double lx = (ax < ex ? ax : ex) - r - mx, hx = (ax > ex ? ax : ex) + r + mx;
    const double ly = (ay < ey ? ay : ey) - r - my, hy = (ay > ey ? ay : ey) + r + my;
    const double lz = (az < ez ? az : ez) - r - mz, hz = (az > ez ? az : ez) + r + mz;
    if (cx[3] < lx || cx[0] > hx || cy[3] < ly || cy[0] > hy || cz[3] < lz || cz[0] > hz) return 0ull;
  }
  const double rr = EP_MUL(r, r);
  // closest_t: diff = b - a, diff_squared (collision_primitives.h:34-44)
  const double dfx = EP_SUB(ex, ax), dfy = EP_SUB(ey, ay), dfz = EP_SUB(ez, az);
  const double d2 = ep_dot(dfx, dfy, dfz, dfx, dfy, dfz);
  const double eps = 2.220446049250313e-16;
  const bool degenerate = d2 <= eps * eps;
  uint64_t bits = 0ull;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 4; k++) {
        const double px = cx[i], py = cy[j], pz = cz[k];
        double qx = ax, qy = ay, qz = az;  // the point the sphere test is made around
        if (capsule) {
          double t = degenerate ? 0.0
                                : EP_DIV(ep_dot(dfx, dfy, dfz, EP_SUB(px, ax), EP_SUB(py, ay), EP_SUB(pz, az)), d2);
          t = (t < 1.0) ? t : 1.0;   // std::min(1.0, t)
          t = (t > 0.0) ? t : 0.0;   // std::max(0.0, .)
          qx = EP_ADD(ax, EP_MUL(dfx, t));  // interpolate: a + (b - a) * t
          qy = EP_ADD(ay, EP_MUL(dfy, t));
          qz = EP_ADD(az, EP_MUL(dfz, t));
        }
        const double sx = EP_SUB(qx, px), sy = EP_SUB(qy, py), sz = EP_SUB(qz, pz);
        if (ep_dot(sx, sy, sz, sx, sy, sz) <= rr) bits |= 1ull << (i * 16 + j * 4 + k);
      }
  return bits;
}

// add_point (VoxelOctree.cpp:319-323): false if p is outside the inclusive limits, else its cell
// (nearest_cell: truncation of (p - min) / d, clamped into the grid)
EP_HD bool ep_point_cell(const double *lo, const double *hi, const double *d, int Ng, const double *p, int *cell) {
  for (int a = 0; a < 3; a++)
    if (!(lo[a] <= p[a] && p[a] <= hi[a])) return false;
  for (int a = 0; a < 3; a++) {
    int i = (int)EP_DIV(EP_SUB(p[a], lo[a]), d[a]);
    i = i < 0 ? 0 : i;
    cell[a] = i > Ng - 1 ? Ng - 1 : i;
  }
  return true;
}
