// roadmapIk of the C++ host mirror (interactive-rate-tendons_b200/host/irt_host.hpp), every branch, against the
// reference's loop (motion-planning/VoxelCachedLazyPRM.cpp:3095-3565) restated here the way the reference runs it:
// one neighbour, one IK, one vertex check, one voxelize_until_invalid at a time, with its early returns, temporary
// vertices and removals, over plain vectors and the CPU oracle (SeqPlanner).  The mirror batches all of that; the
// result (which neighbour, controls, tip, error), the vertices removed and the vertices / edges added with their
// validity must be the same.  Host logic only: linked against the test-only stand-in of the C ABI
// (abi_standin_over_oracle.cpp); built and run by tests/test_abi_and_host.py.  The oracle is the checker only.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <optional>
#include <random>
#include <set>
#include <string>
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

using PRM = motion_planning::VoxelCachedLazyPRM;
using State = std::vector<double>;
using Point = collision::Point;

struct SeqPlanner {
  const tendon::TendonRobot &robot;
  const orc_robot &orb;
  const orc_grid &og;
  const orc_octree *oenv;
  const orc_space &osp;
  const PRM &prm;                 // distance(), getRange() only
  const tip_control::IkSolver &solver;
  double delta;
  std::vector<State> states;
  std::vector<char> alive;
  std::vector<Point> tips;
  std::map<std::pair<size_t, size_t>, unsigned> edges;   // (min, max) -> validity

  struct Partial { bool fully; State last_valid; Point tip; };
  struct Res { size_t i, nbv; State fin; Point tip; double err; bool ok; std::vector<long> nearest; std::vector<Partial> pes; };
  struct Out { std::string kind; size_t i; State controls; Point tip; double error; long source = -1; long vertex = -1; };

  Point fk_tip(const State &x) const {
    std::vector<double> t(512), p(512 * 3);
    orc_fk_out fo;
    const int n = orc_shape(&orb, x.data(), 512, t.data(), p.data(), nullptr, &fo);
    return {p[3 * (n - 1)], p[3 * (n - 1) + 1], p[3 * (n - 1) + 2]};
  }
  bool state_valid(const State &x) const {
    std::vector<double> t(512), p(512 * 3);
    orc_fk_out fo;
    const int n = orc_shape(&orb, x.data(), 512, t.data(), p.data(), nullptr, &fo);
    if (orc_validity_flags(&orb, x.data(), &fo, p.data()) != 0) return false;
    orc_octree *ov = orc_octree_new(&og);
    orc_voxelize_shape(&og, p.data(), n, ov);
    const bool hit = orc_octree_collides(oenv, ov) == 1;
    orc_octree_free(ov);
    return !hit;
  }
  static double norm(const Point &a, const Point &b) {
    double e2 = 0.0;
    for (int c = 0; c < 3; c++) e2 += (a[c] - b[c]) * (a[c] - b[c]);
    return std::sqrt(e2);
  }
  Partial until_invalid(const State &a, const State &b) const {
    orc_octree *ov = orc_octree_new(&og);
    orc_edge_out info;
    orc_voxelize_edge(&orb, &og, &osp, a.data(), b.data(), oenv, ov, &info);
    orc_octree_free(ov);
    State lv(info.last_valid, info.last_valid + a.size());
    return {info.is_fully_valid != 0, lv, fk_tip(lv)};
  }
  bool edge_valid(const State &a, const State &b) const {   // computeEdgeValidity: voxelizeEdge + collides
    orc_octree *ov = orc_octree_new(&og);
    orc_edge_out info;
    orc_voxelize_edge(&orb, &og, &osp, a.data(), b.data(), nullptr, ov, &info);
    const bool ok = info.is_fully_valid && orc_octree_collides(oenv, ov) != 1;
    orc_octree_free(ov);
    return ok;
  }
  long find(const State &x) const {
    for (size_t v = 0; v < states.size(); v++) {
      if (!alive[v]) continue;
      bool same = true;
      for (size_t c = 0; c < x.size(); c++) same = same && std::fabs(states[v][c] - x[c]) <= 2 * 2.220446049250313e-16;
      if (same) return (long)v;
    }
    return -1;
  }
  std::pair<size_t, bool> add(const State &x) {   // addMilestone(state, false, &was_added)
    const long v = find(x);
    if (v >= 0) return {(size_t)v, false};
    states.push_back(x); alive.push_back(1); tips.push_back({0, 0, 0});
    return {states.size() - 1, true};
  }
  void remove(size_t v) {
    alive[v] = 0;
    for (auto it = edges.begin(); it != edges.end();)
      it = (it->first.first == v || it->first.second == v) ? edges.erase(it) : std::next(it);
  }
  size_t degree(size_t v) const {
    size_t c = 0;
    for (auto &e : edges) c += e.first.first == v || e.first.second == v;
    return c;
  }
  std::vector<size_t> conn(size_t v, bool skip_self) const {   // KBoundedStrategy over nn_
    std::vector<std::pair<double, size_t>> d;
    for (size_t u = 0; u < states.size(); u++)
      if (alive[u] && !(skip_self && u == v)) d.emplace_back(prm.distance(states[v], states[u]), u);
    std::sort(d.begin(), d.end());
    std::vector<size_t> out;
    for (size_t j = 0; j < 5 && j < d.size() && d[j].first <= prm.getRange(); j++) out.push_back(d[j].second);
    return out;
  }
  std::vector<size_t> nearest_of(size_t vertex, size_t ik_nb, bool accurate) const {
    if (!accurate) return {ik_nb};
    auto out = conn(vertex, false);
    if (std::find(out.begin(), out.end(), ik_nb) == out.end()) out.push_back(ik_nb);
    return out;
  }
  std::optional<Out> run(const Point &request, double tol, size_t k, bool auto_add, bool accurate, bool lazy_add) {
    std::vector<size_t> nbs;
    for (;;) {
      std::vector<std::pair<double, size_t>> d;
      for (size_t v = 0; v < states.size(); v++) if (alive[v]) d.emplace_back(norm(tips[v], request), v);
      std::sort(d.begin(), d.end());
      nbs.clear();
      for (size_t j = 0; j < k && j < d.size(); j++) nbs.push_back(d[j].second);
      bool bad = false;
      for (size_t v : nbs) if (!state_valid(states[v])) { remove(v); bad = true; }
      if (!bad) break;
    }
    auto fk1 = [&](const State &st) {
      std::vector<Point> tp;
      auto J = tip_control::Jacobian_batch(robot, delta, {st}, &tp, IRT_JAC_LEVMAR_CENTRAL);
      return tip_control::LockstepFk::Eval{tp[0], J[0]};
    };
    std::vector<Res> res;
    for (size_t i = 0; i < nbs.size(); i++) {
      Res r;
      r.i = i; r.nbv = nbs[i];
      r.fin = solver(states[nbs[i]], request, fk1);
      r.tip = fk_tip(r.fin); r.err = norm(r.tip, request); r.ok = state_valid(r.fin);
      if (r.ok && r.err < tol && !auto_add) return Out{"accepted", i, r.fin, r.tip, r.err};
      if (r.err < tol && auto_add) {
        auto [vertex, was_added] = add(r.fin);
        if (degree(vertex) > 0) return Out{"already", i, r.fin, r.tip, r.err};
        for (size_t src : nearest_of(vertex, r.nbv, accurate)) {
          if (!state_valid(states[src])) { if (src != vertex) remove(src); continue; }
          if (src == vertex) { tips[vertex] = r.tip; return Out{"self", i, r.fin, r.tip, r.err, -1, (long)vertex}; }
          r.nearest.push_back((long)src);
          r.pes.push_back(until_invalid(states[src], r.fin));
          if (r.pes.back().fully) {
            edges[std::minmax(src, vertex)] = 1;
            tips[vertex] = r.tip;
            return Out{"connected", i, r.fin, r.tip, r.err, (long)src, (long)vertex};
          }
        }
        if (was_added) remove(vertex);
      }
      res.push_back(r);
    }
    if (!auto_add) {
      const Res *best = nullptr;
      for (auto &r : res) if (r.ok && (!best || r.err < best->err)) best = &r;
      if (best) return Out{"closest_valid", best->i, best->fin, best->tip, best->err};
      std::optional<Out> out;
      for (auto &r : res) {
        auto [vertex, was_added] = add(r.fin);
        for (size_t src : nearest_of(vertex, r.nbv, accurate)) {
          if (!state_valid(states[src])) { if (src != vertex) remove(src); continue; }
          Partial pe = until_invalid(states[src], r.fin);
          const double e = norm(pe.tip, request);
          if (!out || e < out->error) out = Out{"stepped_back", r.i, pe.last_valid, pe.tip, e, (long)src};
        }
        if (was_added) remove(vertex);
      }
      return out;
    }
    std::optional<Out> out;
    Point res_tip{0, 0, 0};
    for (auto &r : res) {
      if (r.nearest.empty()) {
        auto [vertex, was_added] = add(r.fin);
        if (!was_added) continue;
        for (size_t src : nearest_of(vertex, r.nbv, accurate)) {
          if (!state_valid(states[src])) { if (src != vertex) remove(src); continue; }
          r.nearest.push_back(src == vertex ? -1L : (long)src);
          r.pes.push_back(until_invalid(states[src], r.fin));
        }
        remove(vertex);
      }
      for (size_t j = 0; j < r.nearest.size(); j++) {
        const double e = norm(r.pes[j].tip, request);
        if (!out || e < out->error) { out = Out{"fallback", r.i, r.pes[j].last_valid, r.pes[j].tip, e, r.nearest[j]}; res_tip = r.tip; }
      }
    }
    if (!out) return out;
    long vertex = find(out->controls);
    if (vertex < 0) {
      states.push_back(out->controls); alive.push_back(1); tips.push_back(res_tip);
      vertex = (long)states.size() - 1;
      for (size_t n : conn((size_t)vertex, true)) edges.emplace(std::minmax((size_t)vertex, n), 0u);   // not yet in nn_
      out->vertex = vertex;
    }
    if (out->source >= 0 && vertex != out->source) {
      edges[std::minmax((size_t)out->source, (size_t)vertex)] = 1;
      if (!lazy_add)
        for (auto it = edges.begin(); it != edges.end();) {
          const bool mine = it->first.first == (size_t)vertex || it->first.second == (size_t)vertex;
          if (mine && it->second != 1) {
            const size_t other = it->first.first == (size_t)vertex ? it->first.second : it->first.first;
            if (edge_valid(states[(size_t)vertex], states[other])) { it->second = 1; ++it; }
            else it = edges.erase(it);
          } else ++it;
        }
    }
    return out;
  }
};

static orc_robot to_orc(const tendon::TendonRobot &rb) {
  irt_robot_desc d = rb.desc();
  orc_robot o;
  static_assert(sizeof(orc_robot) == sizeof(irt_robot_desc), "POD mirrors must match");
  std::memcpy(&o, &d, sizeof(o));
  return o;
}

int main() {
  tendon::TendonRobot robot;   // robot B: 6 helical tendons, rotation + retraction
  robot.specs.dL = 0.003;
  robot.enable_retraction = true;
  robot.enable_rotation = true;
  for (int k = 0; k < 6; k++) {
    tendon::TendonSpecs t;
    t.C = {k * M_PI / 3, (k % 2 ? -1.0 : 1.0) * 2 * M_PI / robot.specs.L};
    t.D = {0.01};
    robot.tendons.push_back(t);
  }
  orc_robot orb = to_orc(robot);
  collision::VoxelOctree env_vox(128);
  env_vox.set_xlim(-0.21, 0.21); env_vox.set_ylim(-0.21, 0.21); env_vox.set_zlim(-0.21, 0.21);
  orc_grid og;
  std::memset(&og, 0, sizeof(og));
  og.Ng = 128;
  for (int a = 0; a < 3; a++) { og.lim[2 * a] = -0.21; og.lim[2 * a + 1] = 0.21; }
  og.inv_rot[0] = og.inv_rot[4] = og.inv_rot[8] = 1;
  orc_octree *oenv = orc_octree_new(&og);
  const double centres[2][3] = {{0.05, 0.0, 0.12}, {-0.04, 0.03, 0.14}};
  const double radii[2] = {0.03, 0.025};
  for (int q = 0; q < 2; q++) orc_octree_add_sphere(oenv, centres[q], radii[q]);
  {
    std::vector<uint8_t> xyz(3 * 8192);
    std::vector<uint64_t> bits(8192);
    const int64_t nb = orc_octree_export(oenv, 8192, xyz.data(), bits.data());
    for (int64_t i = 0; i < nb; i++) env_vox.set_block(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], bits[i]);
  }
  orc_space osp{0.02, 0.01, 0.0001};
  motion_planning::VoxelEnvironment venv;

  const double Lr = robot.specs.L, delta = 1e-6;
  tip_control::IkSolver dls = [&](const State &start, const Point &req,
                                  const std::function<tip_control::LockstepFk::Eval(const State &)> &fk) {
    State x = start;   // damped least squares with box clamping (stands in for ikController_)
    for (int it = 0; it < 25; it++) {
      auto ev = fk(x);
      double e[3] = {req[0] - ev.tip[0], req[1] - ev.tip[1], req[2] - ev.tip[2]};
      if (std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) < 1e-7) break;
      double A[3][3];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          A[r][c] = (r == c) ? 1e-8 : 0.0;
          for (int j = 0; j < 8; j++) A[r][c] += ev.J[r * 8 + j] * ev.J[c * 8 + j];
        }
      const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                         A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
      double y[3];
      for (int c = 0; c < 3; c++) {   // Cramer
        double B[3][3];
        for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) B[r][q] = (q == c) ? e[r] : A[r][q];
        y[c] = (B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0]) +
                B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0])) / det;
      }
      for (int j = 0; j < 8; j++) {
        x[j] += ev.J[0 * 8 + j] * y[0] + ev.J[1 * 8 + j] * y[1] + ev.J[2 * 8 + j] * y[2];
        const double lo = (j == 6) ? -M_PI : 0.0, hi = (j < 6) ? 20.0 : (j == 6 ? M_PI : Lr);
        x[j] = std::min(hi, std::max(lo, x[j]));
      }
    }
    return x;
  };

  std::mt19937_64 gen(20220802);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::normal_distribution<double> Nrm(0.0, 1.0);
  auto fresh = [&](PRM &prm) {
    prm.setSeed(11);
    prm.createRoadmap(90, PRM::VoxelizeVertices);
    prm.precomputeVertexVoxelCache();
  };
  std::vector<Point> requests;
  {
    PRM base(robot, venv, env_vox);
    fresh(base);
    for (int q = 0; q < 5; q++) {   // reachable tips near roadmap vertices
      State goal = base.states()[gen() % 90];
      for (int j = 0; j < 6; j++) goal[j] = std::min(20.0, std::max(0.0, goal[j] + Nrm(gen)));
      goal[7] = std::min(Lr, std::max(0.0, goal[7] + 0.004 * Nrm(gen)));
      requests.push_back(robot.forward_kinematics(goal).back());
    }
  }
  requests.push_back({0.05, 0.0, 0.12});      // the centre of an obstacle: every result collides
  requests.push_back({-0.04, 0.03, 0.14});
  requests.push_back({0.12, 0.12, 0.19});     // out of reach: no result within tolerance

  std::set<std::string> seen;
  long lazy_edges = 0, eager_edges = 0;
  for (int accurate = 0; accurate < 2; accurate++)
    for (auto &request : requests)
      for (int variant = 0; variant < 3; variant++) {
        const bool auto_add = variant > 0, lazy_add = variant == 2;
        PRM prm(robot, venv, env_vox);
        fresh(prm);
        SeqPlanner seq{robot, orb, og, oenv, osp, prm, dls, delta, prm.states(), {}, {}, {}};
        seq.alive.assign(seq.states.size(), 1);
        for (size_t v = 0; v < seq.states.size(); v++)
          seq.tips.push_back({prm.tipPositions()[3 * v], prm.tipPositions()[3 * v + 1], prm.tipPositions()[3 * v + 2]});
        for (size_t e = 0; e < prm.edges().size(); e++)
          seq.edges[std::minmax(prm.edges()[e].first, prm.edges()[e].second)] = prm.edgeValidity()[e];
        const size_t nv = prm.states().size();
        auto want = seq.run(request, 1e-4, 4, auto_add, accurate, lazy_add);
        const unsigned opt = (auto_add ? PRM::RMAP_IK_AUTO_ADD : 0u) | (accurate ? PRM::RMAP_IK_ACCURATE : 0u) |
                             (lazy_add ? PRM::RMAP_IK_LAZY_ADD : 0u);
        auto got = prm.roadmapIk(request, 1e-4, 4, dls, IRT_JAC_LEVMAR_CENTRAL, delta, opt);
        CHECK(want.has_value() == got.has_value());
        if (!want || !got) { seen.insert("none"); continue; }
        seen.insert(want->kind);
        CHECK(got->index == want->i);
        CHECK(got->controls == want->controls);
        for (int c = 0; c < 3; c++) CHECK(got->tip_position[c] == want->tip[c]);
        CHECK(got->error == want->error);
        CHECK(got->stepped_back == (want->kind == "stepped_back" || want->kind == "fallback"));
        if (want->source >= 0) CHECK(got->source == want->source);
        // the graph afterwards: the living vertices in order, the edges between them with their validity
        std::vector<size_t> live_seq, live_got;
        for (size_t v = 0; v < seq.states.size(); v++) if (seq.alive[v]) live_seq.push_back(v);
        for (size_t v = 0; v < prm.states().size(); v++)
          if (v >= prm.removedVertices().size() || !prm.removedVertices()[v]) live_got.push_back(v);
        CHECK(live_seq.size() == live_got.size());
        if (live_seq.size() != live_got.size()) {
          std::printf("  accurate %d variant %d kind %s: %zu living vertices, want %zu (of %zu / %zu)\n", accurate, variant,
                      want->kind.c_str(), live_got.size(), live_seq.size(), prm.states().size(), seq.states.size());
          continue;
        }
        std::map<size_t, size_t> remap;
        for (size_t j = 0; j < live_seq.size(); j++) {
          remap[live_seq[j]] = live_got[j];
          CHECK(seq.states[live_seq[j]] == prm.states()[live_got[j]]);
          if (live_seq[j] < nv) CHECK(live_seq[j] == live_got[j]);   // the same old vertices were removed
        }
        std::map<std::pair<size_t, size_t>, unsigned> have, expect;
        for (size_t e = 0; e < prm.edges().size(); e++) {
          const auto &ed = prm.edges()[e];
          if (prm.removedEdges()[e] || prm.removedVertices()[ed.first] || prm.removedVertices()[ed.second]) continue;
          have[std::minmax(ed.first, ed.second)] = prm.edgeValidity()[e];
        }
        for (auto &ed : seq.edges) expect[std::minmax(remap.at(ed.first.first), remap.at(ed.first.second))] = ed.second;
        CHECK(have == expect);
        if (want->vertex >= 0 && want->kind != "self") {
          CHECK(got->added_vertex == (long)remap.at((size_t)want->vertex));
          long mine = 0;
          for (auto &ed : have) mine += ed.first.first == (size_t)got->added_vertex || ed.first.second == (size_t)got->added_vertex;
          if (want->kind == "fallback") (lazy_add ? lazy_edges : eager_edges) += mine;
          for (int c = 0; c < 3; c++)
            CHECK(prm.tipPositions()[3 * (size_t)got->added_vertex + c] == seq.tips[(size_t)want->vertex][c]);
        }
      }
  // ---- chainedPlan: the milestone loop of apps/roadmap_chained_plan.cpp:535-679 ----------------------------------
  {
    PRM prm(robot, venv, env_vox);
    fresh(prm);
    prm.precomputeEdgeVoxelCache();
    prm.precomputeValidity();                       // the tick's two sweeps; the chain runs on their tables
    const size_t nv0 = prm.states().size(), ne0 = prm.edges().size(), sweeps0 = prm.sweepCount();
    const size_t cache_v = prm.vertexVoxelCacheSize(), cache_e = prm.edgeVoxelCacheSize();
    State start = prm.states()[3];
    for (int j = 0; j < 6; j++) start[j] = std::min(20.0, start[j] + 0.05);   // off the roadmap: becomes a milestone
    std::vector<Point> reqs(requests.begin(), requests.begin() + 5);
    auto chain = prm.chainedPlan(start, reqs, 1e-4, 4, dls, PRM::RMAP_IK_AUTO_ADD, IRT_JAC_LEVMAR_CENTRAL, delta);
    CHECK(chain.size() == reqs.size());
    SeqPlanner orc_check{robot, orb, og, oenv, osp, prm, dls, delta, {}, {}, {}, {}};
    // truth of the roadmap as it is now, item by item by the oracle
    std::vector<char> v_ok(prm.states().size()), e_ok(prm.edges().size());
    for (size_t v = 0; v < v_ok.size(); v++) v_ok[v] = orc_check.state_valid(prm.states()[v]);
    for (size_t e = 0; e < e_ok.size(); e++)
      e_ok[e] = !prm.removedEdges()[e] && orc_check.edge_valid(prm.states()[prm.edges()[e].first], prm.states()[prm.edges()[e].second]);
    // Dijkstra over the valid sub-graph (ends always kept) of the roadmap as it was when the search ran
    auto shortest = [&](size_t a, size_t b, size_t n, size_t n_edges) {
      std::vector<double> dist(n, 1e300);
      std::vector<char> done(n, 0);
      dist[a] = 0.0;
      for (;;) {
        size_t u = n;
        for (size_t v = 0; v < n; v++) if (!done[v] && dist[v] < 1e300 && (u == n || dist[v] < dist[u])) u = v;
        if (u == n || u == b) break;
        done[u] = 1;
        for (size_t e = 0; e < n_edges; e++) {
          if (!e_ok[e]) continue;
          const auto &ed = prm.edges()[e];
          if (ed.first != u && ed.second != u) continue;
          const size_t w = ed.first == u ? ed.second : ed.first;
          if (!(v_ok[w] || w == a || w == b) || prm.removedVertices()[w]) continue;
          dist[w] = std::min(dist[w], dist[u] + prm.distance(prm.states()[u], prm.states()[w]));
        }
      }
      return dist[b];
    };
    State current = start;
    int exact = 0;
    for (auto &m : chain) {
      CHECK(prm.states()[m.start_vertex] == current && prm.states()[m.goal_vertex] == m.ik.controls);
      CHECK(v_ok[m.goal_vertex]);
      const double want = shortest(m.start_vertex, m.goal_vertex, m.n_vertices, m.n_edges);
      if (m.exact) {
        exact++;
        CHECK(m.path.front() == m.start_vertex && m.path.back() == m.goal_vertex && m.plan.size() == m.path.size());
        double cost = 0.0;
        for (size_t q = 1; q < m.path.size(); q++) {
          const long e = prm.edgeIndex(m.path[q - 1], m.path[q]);
          CHECK(e >= 0 && e_ok[(size_t)e]);
          if (q + 1 < m.path.size()) CHECK(v_ok[m.path[q]]);
          cost += prm.distance(prm.states()[m.path[q - 1]], prm.states()[m.path[q]]);
        }
        CHECK(std::fabs(cost - want) <= 1e-9 * std::max(1.0, want));
        CHECK(std::fabs(m.tip_error - m.ik.error) < 1e-12);
      } else {
        CHECK(m.path.empty() && m.plan.size() == 1 && want >= 1e300);
      }
      current = m.plan.back();
    }
    CHECK(exact >= 3);
    CHECK(prm.states().size() > nv0 && prm.edges().size() > ne0);         // milestones and lazy connections joined
    CHECK(prm.sweepCount() == sweeps0);                                   // no further sweep during the chain
    CHECK(prm.singleCheckCount() > 0 && prm.singleCheckCount() <= (prm.states().size() - nv0) + (prm.edges().size() - ne0));
    // an environment change: the next sweeps cover the newcomers from a scratch batch, the caches stay
    prm.setEnvironment(env_vox.empty_copy());
    prm.restoreRemoved();
    prm.precomputeValidity();
    CHECK(prm.vertexVoxelCacheSize() == cache_v && prm.edgeVoxelCacheSize() == cache_e);
    CHECK(prm.vertexValidity().size() == prm.states().size() && prm.edgeValidity().size() == prm.edges().size());
    orc_octree *empty = orc_octree_new(&og);
    SeqPlanner free_check{robot, orb, og, empty, osp, prm, dls, delta, {}, {}, {}, {}};
    for (size_t v = nv0; v < prm.states().size(); v++)
      CHECK((prm.vertexValidity()[v] == 1u) == free_check.state_valid(prm.states()[v]));
    for (size_t e = ne0; e < prm.edges().size(); e++)
      CHECK((prm.edgeValidity()[e] == 1u) == free_check.edge_valid(prm.states()[prm.edges()[e].first], prm.states()[prm.edges()[e].second]));
    orc_octree_free(empty);
    prm.setMaxUncached(0);                          // too many newcomers: the next sweep rebuilds the caches
    prm.clearValidity();
    prm.precomputeValidity();
    CHECK(prm.vertexVoxelCacheSize() == prm.states().size() && prm.edgeVoxelCacheSize() == prm.edges().size());
    std::printf("chainedPlan: %d of %zu milestones reached exactly, %zu newcomers checked one at a time, %zu vertices / %zu edges joined\n",
                exact, chain.size(), prm.singleCheckCount(), prm.states().size() - nv0, prm.edges().size() - ne0);
  }
  for (const char *kind : {"accepted", "closest_valid", "stepped_back", "connected", "self", "fallback"})
    if (!seen.count(kind)) { std::printf("branch never reached: %s\n", kind); failures++; }
  CHECK(eager_edges < lazy_edges);   // validation removed some of the lazily connected edges
  std::printf("roadmapIk branches:");
  for (auto &s : seen) std::printf(" %s", s.c_str());
  std::printf("; edges of the added vertex: %ld lazy, %ld validated\n", lazy_edges, eager_edges);
  orc_octree_free(oenv);
  std::printf(failures ? "FAILED (%d)\n" : "roadmap ik ok\n", failures);
  return failures ? 1 : 0;
}
