// The reference's self-collision test, compiled from its own text (test infrastructure only).
//
// collision/collision.cpp needs FCL, but collides_self (collision.cpp:6-46) and what it calls do not.
// oracle/Makefile cuts, by anchors, into oracle/_ref/gen/ (deleted after the build):
//   sc_sphere.inc / sc_capsule.inc / sc_capseq.inc   struct Sphere / Capsule / CapsuleSequence from
//        collision/{Sphere,Capsule,CapsuleSequence}.h, each truncated at its first cpptoml / fcl member
//        (fields, operator==, operator[], at, size, interpolate, closest_t remain)
//   sc_collides.inc   the inline collides() overloads for Point / Sphere / Capsule, collision.hxx:55-108
//   sc_self.inc       collides_self(const CapsuleSequence&), collision.cpp:6-46
// closest_st_segment comes from the unmodified collision_primitives.cpp; Point arithmetic through the
// Eigen stand-in (pins the reference's expressions, not Eigen's rounding).
#include <collision/Point.h>
#include <collision/collision_primitives.h>

#include <algorithm>
#include <cmath>
#include <iostream>
#include <memory>
#include <tuple>
#include <utility>
#include <vector>

namespace collision {
#include "sc_sphere.inc"
#include "sc_capsule.inc"
#include "sc_capseq.inc"
#include "sc_collides.inc"
#include "sc_self.inc"
}  // namespace collision

extern "C" {

int scref_collides_self(const double *p, int n, double r) {
  collision::CapsuleSequence seq;
  seq.r = r;
  for (int i = 0; i < n; i++) seq.points.emplace_back(p + 3 * i);
  return collision::collides_self(seq) ? 1 : 0;
}

int scref_capsules_collide(const double *a0, const double *a1, double ra, const double *b0,
                           const double *b1, double rb) {
  return collision::collides(collision::Capsule{collision::Point(a0), collision::Point(a1), ra},
                             collision::Capsule{collision::Point(b0), collision::Point(b1), rb}) ? 1 : 0;
}

}  // extern "C"
