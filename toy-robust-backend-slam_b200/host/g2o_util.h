// g2o_util.h — g2o reader, Vertigo-style outlier injection and text writers of the drop-in
// `main`, dependency-free (the reference needs boost, absent here).
//
// Behaviour follows the reference's ReadG2O (DCS-ceres/include/g2o_util.h):
//   * reader            :20-89   tags VERTEX_SE2|VERTEX2 and EDGE_SE2|EDGE2, tokens split on
//                                single spaces with compression, vertex ids index nNodes directly
//                                (ids must be 0..N-1 in file order, vertices before edges),
//                                edge class = |a-b| < 5 ? odometry : closure (:68)
//   * writers           :93-112, :179-186   "index x y theta" / "a b type", 6 significant digits
//   * add_random_C      :151-171 libc rand(): a, b, (a==b -> b=(b+1)%N), then three rand()/RAND_MAX
//                                integer divisions (always 0 unless rand()==RAND_MAX)
// Same public member names as the reference so main.cpp reads the same.  New: flatten()
// (AoS -> SoA for the C-ABI) and scatter_poses() (write-back of the solved state).
#ifndef DCS_B200_G2O_UTIL_H
#define DCS_B200_G2O_UTIL_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <string>
#include <vector>

#include "graph.h"

// Flat structure-of-arrays view handed to dcs_create (include/dcs_b200.h: dcs_graph).
struct FlatGraph {
  std::vector<double> pose_xyt;   // N x 3
  std::vector<int32_t> edge_a, edge_b;
  std::vector<double> meas_xyt;   // E x 3
  std::vector<uint8_t> kind;      // order: odometry, closure, bogus (reference main.cpp:95-150)
};

class ReadG2O {
 public:
  ReadG2O() {}
  explicit ReadG2O(const std::string& fName) { read(fName); }
  ReadG2O(const ReadG2O&) = delete;
  ReadG2O& operator=(const ReadG2O&) = delete;

  // Returns false when the file cannot be opened (the reference silently yields an empty graph).
  bool read(const std::string& fName) {
    FILE* fp = std::fopen(fName.c_str(), "rb");
    if (!fp) return false;
    std::fseek(fp, 0, SEEK_END);
    const long sz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::string buf((size_t)(sz > 0 ? sz : 0), '\0');
    const size_t got = sz > 0 ? std::fread(&buf[0], 1, (size_t)sz, fp) : 0;
    std::fclose(fp);
    buf.resize(got);
    parse(buf);
    return true;
  }

  // Parses g2o text already in memory (also used by the tests).
  void parse(const std::string& text) {
    const char* s = text.data();
    const char* end = s + text.size();
    const char* tok[16];
    while (s < end) {
      const char* eol = (const char*)std::memchr(s, '\n', (size_t)(end - s));
      if (!eol) eol = end;
      // split on ' ' with compression; a leading space yields an empty first token
      int nt = 0;
      const char* q = s;
      if (q < eol && *q == ' ') { tok[nt++] = q; while (q < eol && *q == ' ') ++q; }
      while (q < eol && nt < 16) {
        tok[nt++] = q;
        while (q < eol && *q != ' ') ++q;
        while (q < eol && *q == ' ') ++q;
      }
      if (nt >= 5 && (tag_is(tok[0], eol, "VERTEX_SE2") || tag_is(tok[0], eol, "VERTEX2"))) {
        const int idx = (int)std::strtol(tok[1], nullptr, 10);
        add_node(idx, std::strtod(tok[2], nullptr), std::strtod(tok[3], nullptr), std::strtod(tok[4], nullptr));
      } else if (nt >= 12 && (tag_is(tok[0], eol, "EDGE_SE2") || tag_is(tok[0], eol, "EDGE2"))) {
        const int a = (int)std::strtol(tok[1], nullptr, 10);
        const int b = (int)std::strtol(tok[2], nullptr, 10);
        if (a >= 0 && b >= 0 && a < (int)nNodes.size() && b < (int)nNodes.size()) {
          const int type = (std::abs(a - b) < 5) ? ODOMETRY_EDGE : CLOSURE_EDGE;
          Edge* e = new_edge(nNodes[a], nNodes[b], type);
          e->setEdgePose(std::strtod(tok[3], nullptr), std::strtod(tok[4], nullptr), std::strtod(tok[5], nullptr));
          e->setInformationMatrix(std::strtod(tok[6], nullptr), std::strtod(tok[7], nullptr), std::strtod(tok[8], nullptr),
                                  std::strtod(tok[9], nullptr), std::strtod(tok[10], nullptr), std::strtod(tok[11], nullptr));
          (type == ODOMETRY_EDGE ? nEdgesOdometry : nEdgesClosure).push_back(e);
        }
      }
      s = eol + 1;
    }
  }

  Node* add_node(int index, double x, double y, double theta) {
    pose_arena_.emplace_back();
    double* p = pose_arena_.back().v;
    node_arena_.emplace_back(index, p, x, y, theta);
    node_arena_.back().slot = (int)nNodes.size();
    nNodes.push_back(&node_arena_.back());
    return nNodes.back();
  }

  Edge* new_edge(const Node* a, const Node* b, int type) {
    edge_arena_.emplace_back(a, b, type);
    return &edge_arena_.back();
  }

  // Adding bogus edges as described in the Vertigo paper (reference :151-171).
  void add_random_C(int count, bool print = true) {
    const int MAX = (int)nNodes.size();
    std::cout << "Adding Bogus edges as described in Vertigo paper" << std::endl;
    if (MAX <= 0) return;
    std::string out;
    char line[64];
    for (int i = 0; i < count; ++i) {
      const int a = std::rand() % MAX;
      int b = std::rand() % MAX;
      if (a == b) b = (b + 1) % MAX;  // no self loops (two identical parameter blocks abort Ceres)
      if (print) { std::snprintf(line, sizeof(line), "  %d<--->%d\n", a, b); out += line; }
      Edge* e = new_edge(nNodes[a], nNodes[b], BOGUS_EDGE);
      // Three integer divisions, as the reference wrote them; gcc on x86-64 evaluates the call
      // arguments right to left, so the first draw lands in theta.
      const double th = (double)(std::rand() / RAND_MAX);
      const double y = (double)(std::rand() / RAND_MAX);
      const double x = (double)(std::rand() / RAND_MAX);
      e->setEdgePose(x, y, th);
      e->setInformationMatrix(2.0, 0.0, 0.0, 300.0, 0.0, 300.0);
      nEdgesBogus.push_back(e);
      if (out.size() > (1u << 16)) { std::cout << out; out.clear(); }
    }
    std::cout << out;
  }

  void writePoseGraph_nodes(const std::string& fname) {
    std::cout << "writePoseGraph nodes: " << fname << std::endl;
    FILE* fp = std::fopen(fname.c_str(), "w");
    if (!fp) return;
    for (const Node* n : nNodes) std::fprintf(fp, "%d %g %g %g\n", n->index, n->p[0], n->p[1], n->p[2]);
    std::fclose(fp);
  }

  void writePoseGraph_edges(const std::string& fname) {
    std::cout << "writePoseGraph Edges : " << fname << std::endl;
    FILE* fp = std::fopen(fname.c_str(), "w");
    if (!fp) return;
    write_edges(fp, nEdgesOdometry);
    write_edges(fp, nEdgesClosure);
    write_edges(fp, nEdgesBogus);
    std::fclose(fp);
  }

  // AoS -> SoA in the residual-block order of the reference's main.cpp:95-150.
  void flatten(FlatGraph* g) const {
    const size_t N = nNodes.size();
    const size_t E = nEdgesOdometry.size() + nEdgesClosure.size() + nEdgesBogus.size();
    g->pose_xyt.resize(N * 3);
    for (size_t i = 0; i < N; ++i) std::memcpy(&g->pose_xyt[3 * i], nNodes[i]->p, 3 * sizeof(double));
    g->edge_a.clear(); g->edge_b.clear(); g->meas_xyt.clear(); g->kind.clear();
    g->edge_a.reserve(E); g->edge_b.reserve(E); g->meas_xyt.reserve(3 * E); g->kind.reserve(E);
    const std::vector<Edge*>* groups[3] = {&nEdgesOdometry, &nEdgesClosure, &nEdgesBogus};
    for (const auto* grp : groups)
      for (const Edge* e : *grp) {
        // position in nNodes, not Node::index: the reference hands Ceres the Node::p pointers
        g->edge_a.push_back((int32_t)e->a->slot);
        g->edge_b.push_back((int32_t)e->b->slot);
        g->meas_xyt.push_back(e->x); g->meas_xyt.push_back(e->y); g->meas_xyt.push_back(e->theta);
        g->kind.push_back((uint8_t)e->edge_type);
      }
  }

  // Solved state -> Node::p, in place (what ceres::Solve does to the user's parameter blocks).
  void scatter_poses(const double* pose_xyt) {
    for (size_t i = 0; i < nNodes.size(); ++i) std::memcpy(nNodes[i]->p, pose_xyt + 3 * i, 3 * sizeof(double));
  }

  std::vector<Node*> nNodes;
  std::vector<Edge*> nEdgesOdometry;
  std::vector<Edge*> nEdgesClosure;
  std::vector<Edge*> nEdgesBogus;

 private:
  struct P3 { double v[3]; };
  // deques: stable addresses while growing
  std::deque<P3> pose_arena_;
  std::deque<Node> node_arena_;
  std::deque<Edge> edge_arena_;

  static bool tag_is(const char* t, const char* eol, const char* tag) {
    const size_t n = std::strlen(tag);
    return (size_t)(eol - t) >= n && std::memcmp(t, tag, n) == 0 && (t + n == eol || t[n] == ' ');
  }

  static void write_edges(FILE* fp, const std::vector<Edge*>& vec) {
    for (const Edge* e : vec) std::fprintf(fp, "%d %d %d\n", e->a->index, e->b->index, e->edge_type);
  }
};

#endif
