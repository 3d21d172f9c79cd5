// synth.h — synthetic 2D Manhattan-world pose graphs for the benchmark configs of BASELINE.json
// ("synthetic 2D Manhattan grid 1M poses / 4M edges, 10% outlier loops").  The reference has no
// generator; the output is a graph its reader would accept (ids 0..N-1, vertices first,
// 12-token EDGE_SE2 lines; DCS-ceres/include/g2o_util.h:33-87) so the same host path is used.
//
//   world      : walk on a W x W integer grid, W = ceil(sqrt(N)); unit steps; at each step turn
//                +90 / -90 degrees with probability 0.2 each; reflect at the walls
//   odometry   : (i, i+1), ground-truth relative pose + N(0, 0.02) on x,y and N(0, 0.005) on theta
//   loops      : (j, i), j < i-5, j an earlier visit of i's cell or of one of its 8 neighbours
//                (most recent first).  A per-pose cap c is chosen so that the total is exactly
//                n_loops: every pose takes min(avail, c-1), the first poses with avail >= c take
//                one more.  Same noise model.
//   vertices   : dead-reckoned from the noisy odometry (the usual initial guess)
//   RNG        : splitmix64 -> Box-Muller, explicit (no <random> distributions), so graphs are
//                bit-reproducible across libstdc++ versions.
#ifndef DCS_B200_SYNTH_H
#define DCS_B200_SYNTH_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "g2o_util.h"

namespace synth {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double normal() {
    double u1 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    const double u2 = uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
  }
};

struct Pose { double x, y, th; };

inline Pose between(const Pose& a, const Pose& b) {  // a^-1 * b
  const double c = std::cos(a.th), s = std::sin(a.th);
  const double dx = b.x - a.x, dy = b.y - a.y;
  double dth = b.th - a.th;
  while (dth > M_PI) dth -= 2.0 * M_PI;
  while (dth <= -M_PI) dth += 2.0 * M_PI;
  return Pose{c * dx + s * dy, -s * dx + c * dy, dth};
}
inline Pose compose(const Pose& a, const Pose& d) {  // a * d
  const double c = std::cos(a.th), s = std::sin(a.th);
  return Pose{a.x + c * d.x - s * d.y, a.y + s * d.x + c * d.y, a.th + d.th};
}

static const double kSigmaXY = 0.02, kSigmaTh = 0.005;
static const int kMaxAvail = 8;

// Fills `out` (must be empty).  Returns the number of loop edges actually produced
// (== n_loops unless the walk offers fewer candidates than asked for).
inline int64_t generate_manhattan(int N, int64_t n_loops, uint64_t seed, ReadG2O* out) {
  if (N < 2) return -1;
  Rng rng(seed);
  const int W = (int)std::ceil(std::sqrt((double)N));
  std::vector<int> cx(N), cy(N);
  std::vector<int8_t> hd(N);
  static const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1};
  int x = W / 2, y = W / 2, h = 0;
  cx[0] = x; cy[0] = y; hd[0] = 0;
  for (int i = 1; i < N; ++i) {
    const double u = rng.uniform();
    if (u < 0.2) h = (h + 1) & 3; else if (u < 0.4) h = (h + 3) & 3;
    int nx = x + DX[h], ny = y + DY[h];
    if (nx < 0 || nx >= W || ny < 0 || ny >= W) { h = (h + 2) & 3; nx = x + DX[h]; ny = y + DY[h]; }
    x = nx; y = ny;
    cx[i] = x; cy[i] = y; hd[i] = (int8_t)h;
  }
  auto gt = [&](int i) { double th = hd[i] * (M_PI / 2.0); if (th > M_PI) th -= 2.0 * M_PI; return Pose{(double)cx[i], (double)cy[i], th}; };

  // odometry + dead reckoning
  std::vector<Pose> odo(N - 1);
  Pose cur = gt(0);
  out->add_node(0, cur.x, cur.y, cur.th);
  for (int i = 0; i + 1 < N; ++i) {
    Pose d = between(gt(i), gt(i + 1));
    d.x += kSigmaXY * rng.normal(); d.y += kSigmaXY * rng.normal(); d.th += kSigmaTh * rng.normal();
    odo[i] = d;
    cur = compose(cur, d);
    out->add_node(i + 1, cur.x, cur.y, cur.th);
  }
  for (int i = 0; i + 1 < N; ++i) {
    Edge* e = out->new_edge(out->nNodes[i], out->nNodes[i + 1], ODOMETRY_EDGE);
    e->setEdgePose(odo[i].x, odo[i].y, odo[i].th);
    e->setInformationMatrix(44.721360, 0, 0, 44.721360, 0, 44.721360);
    out->nEdgesOdometry.push_back(e);
  }

  // loop candidates: earlier visits of the same / neighbouring cells, most recent first
  std::vector<int> last((size_t)W * W, -1), prev(N, -1);
  std::vector<int> cand((size_t)N * kMaxAvail, -1);
  std::vector<uint8_t> avail(N, 0);
  for (int i = 0; i < N; ++i) {
    int na = 0;
    int heads[9], nh = 0;
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        const int qx = cx[i] + ox, qy = cy[i] + oy;
        if (qx < 0 || qx >= W || qy < 0 || qy >= W) continue;
        heads[nh++] = last[(size_t)qy * W + qx];
      }
    // k-way merge by recency over the (up to 9) per-cell chains
    while (na < kMaxAvail) {
      int best = -1, bi = -1;
      for (int q = 0; q < nh; ++q) if (heads[q] > best) { best = heads[q]; bi = q; }
      if (best < 0) break;
      heads[bi] = prev[best];
      if (best < i - 5) cand[(size_t)i * kMaxAvail + na++] = best;
    }
    avail[i] = (uint8_t)na;
    prev[i] = last[(size_t)cy[i] * W + cx[i]];
    last[(size_t)cy[i] * W + cx[i]] = i;
  }
  // choose the cap
  int64_t hist[kMaxAvail + 1] = {0};
  for (int i = 0; i < N; ++i) hist[avail[i]]++;
  int cap = 0;
  int64_t total = 0, below = 0;
  for (cap = 1; cap <= kMaxAvail; ++cap) {
    below = total;
    int64_t ge = 0;
    for (int a = cap; a <= kMaxAvail; ++a) ge += hist[a];
    total += ge;  // sum_i min(avail_i, cap)
    if (total >= n_loops) break;
  }
  if (cap > kMaxAvail) { cap = kMaxAvail; below = total; }
  int64_t extra = n_loops - below;  // poses with avail >= cap that take one more
  if (extra < 0) extra = 0;
  int64_t made = 0;
  for (int i = 0; i < N; ++i) {
    int take = avail[i] < cap - 1 ? avail[i] : cap - 1;
    if (avail[i] >= cap && extra > 0) { ++take; --extra; }
    for (int q = 0; q < take; ++q) {
      const int j = cand[(size_t)i * kMaxAvail + q];
      Pose d = between(gt(j), gt(i));
      d.x += kSigmaXY * rng.normal(); d.y += kSigmaXY * rng.normal(); d.th += kSigmaTh * rng.normal();
      Edge* e = out->new_edge(out->nNodes[j], out->nNodes[i], CLOSURE_EDGE);
      e->setEdgePose(d.x, d.y, d.th);
      e->setInformationMatrix(44.721360, 0, 0, 44.721360, 0, 44.721360);
      out->nEdgesClosure.push_back(e);
      ++made;
    }
  }
  return made;
}

// g2o text the (unchanged) reader consumes.  17 significant digits so a round trip is exact.
inline bool write_g2o(const ReadG2O& g, const std::string& path) {
  FILE* fp = std::fopen(path.c_str(), "w");
  if (!fp) return false;
  for (const Node* n : g.nNodes) std::fprintf(fp, "VERTEX_SE2 %d %.17g %.17g %.17g\n", n->index, n->p[0], n->p[1], n->p[2]);
  const std::vector<Edge*>* groups[2] = {&g.nEdgesOdometry, &g.nEdgesClosure};
  for (const auto* grp : groups)
    for (const Edge* e : *grp)
      std::fprintf(fp, "EDGE_SE2 %d %d %.17g %.17g %.17g %g %g %g %g %g %g\n", e->a->index, e->b->index, e->x, e->y,
                   e->theta, e->I11, e->I12, e->I13, e->I22, e->I23, e->I33);
  std::fclose(fp);
  return true;
}

}  // namespace synth
#endif
