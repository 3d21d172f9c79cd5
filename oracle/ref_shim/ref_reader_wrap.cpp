// C wrapper around the REFERENCE's own reader / injector / writers (DCS-ceres/include/g2o_util.h,
// graph.h), compiled from /root/reference where they lie against the boost stand-ins in this
// directory -> oracle/_ref/libdcs_ref_reader.so.  TEST INFRASTRUCTURE ONLY: pins the host reader.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include "g2o_util.h"

struct ref_graph { ReadG2O* g; };

extern "C" {
void* ref_read(const char* path, unsigned seed, int n_bogus) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  ReadG2O* g = nullptr;
  try {
    g = new ReadG2O(std::string(path));
    std::srand(seed);
    g->add_random_C(n_bogus);
  } catch (...) { g = nullptr; }
  std::cout.rdbuf(old);
  return g;
}
void ref_counts(void* h, int* n) {
  ReadG2O* g = (ReadG2O*)h;
  n[0] = (int)g->nNodes.size(); n[1] = (int)g->nEdgesOdometry.size(); n[2] = (int)g->nEdgesClosure.size(); n[3] = (int)g->nEdgesBogus.size();
}
// residual-block order of main.cpp:95-150; endpoints as Node::index
void ref_flatten(void* h, double* pose, int* ea, int* eb, double* meas, unsigned char* kind) {
  ReadG2O* g = (ReadG2O*)h;
  for (size_t i = 0; i < g->nNodes.size(); ++i) std::memcpy(pose + 3 * i, g->nNodes[i]->p, 24);
  size_t k = 0;
  std::vector<Edge*>* grp[3] = {&g->nEdgesOdometry, &g->nEdgesClosure, &g->nEdgesBogus};
  for (auto* v : grp)
    for (Edge* e : *v) {
      ea[k] = e->a->index; eb[k] = e->b->index;
      meas[3 * k] = e->x; meas[3 * k + 1] = e->y; meas[3 * k + 2] = e->theta;
      kind[k] = (unsigned char)e->edge_type;
      ++k;
    }
}
void ref_write(void* h, const char* nodes, const char* edges) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  ReadG2O* g = (ReadG2O*)h;
  g->writePoseGraph_nodes(nodes);
  g->writePoseGraph_edges(edges);
  std::cout.rdbuf(old);
}
// METHOD 2 writer (g2o_util.h:114-148): priors / optimised switch values per loop edge (closure, then bogus)
void ref_write_switches(void* h, const char* path, const double* priors, const double* optimized, int n) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  ReadG2O* g = (ReadG2O*)h;
  std::vector<double> pr(priors, priors + n);
  std::vector<double> vals(optimized, optimized + n);
  std::vector<double*> opt;
  for (int i = 0; i < n; ++i) opt.push_back(&vals[i]);
  g->writePoseGraph_switches(path, pr, opt);
  std::cout.rdbuf(old);
}
}
