// host_capi.cpp — C entry points over the host-side reader / injector / writers / generator
// (libdcs_host.so), so the Python test and bench harness drives the SAME C++ host code the
// drop-in `main` uses.  No CUDA here.
#include <cstdlib>
#include <cstring>
#include <new>
#include <sstream>

#include "g2o_util.h"
#include "synth.h"

struct dcs_host_graph {
  ReadG2O g;
  FlatGraph flat;
  bool flat_valid = false;
};

namespace {
struct CoutSilencer {  // the reader / injector print the reference's stdout lines
  std::streambuf* old;
  std::ostringstream sink;
  explicit CoutSilencer(bool on) : old(nullptr) { if (on) old = std::cout.rdbuf(sink.rdbuf()); }
  ~CoutSilencer() { if (old) std::cout.rdbuf(old); }
};
void ensure_flat(dcs_host_graph* h) {
  if (!h->flat_valid) { h->g.flatten(&h->flat); h->flat_valid = true; }
}
}  // namespace

// Only the entry points below are exported (-fvisibility=hidden); none of them shares a name with a symbol of
// libdcs_b200.so (tests/test_host_cpu.py::test_the_two_libraries_export_disjoint_symbols).
#define DCS_HOST_API __attribute__((visibility("default")))

extern "C" {

DCS_HOST_API dcs_host_graph* dcs_host_read_g2o(const char* path) {
  dcs_host_graph* h = new (std::nothrow) dcs_host_graph();
  if (!h) return nullptr;
  if (!h->g.read(path)) { delete h; return nullptr; }
  return h;
}

DCS_HOST_API dcs_host_graph* dcs_host_parse_g2o(const char* text, int64_t len) {
  dcs_host_graph* h = new (std::nothrow) dcs_host_graph();
  if (!h) return nullptr;
  h->g.parse(std::string(text, (size_t)len));
  return h;
}

DCS_HOST_API dcs_host_graph* dcs_host_synth_manhattan(int32_t n_poses, int64_t n_loops, uint64_t seed, int64_t* made) {
  dcs_host_graph* h = new (std::nothrow) dcs_host_graph();
  if (!h) return nullptr;
  const int64_t m = synth::generate_manhattan(n_poses, n_loops, seed, &h->g);
  if (made) *made = m;
  if (m < 0) { delete h; return nullptr; }
  return h;
}

// srand(seed) + add_random_C(count): the reference seeds with time(0) (main.cpp:43).
DCS_HOST_API void dcs_host_add_random_C(dcs_host_graph* h, int32_t count, uint32_t seed, int32_t quiet) {
  CoutSilencer s(quiet != 0);
  std::srand(seed);
  h->g.add_random_C(count, quiet == 0);
  h->flat_valid = false;
}

DCS_HOST_API void dcs_host_counts(dcs_host_graph* h, int32_t* n_nodes, int32_t* n_odometry, int32_t* n_closure, int32_t* n_bogus) {
  if (n_nodes) *n_nodes = (int32_t)h->g.nNodes.size();
  if (n_odometry) *n_odometry = (int32_t)h->g.nEdgesOdometry.size();
  if (n_closure) *n_closure = (int32_t)h->g.nEdgesClosure.size();
  if (n_bogus) *n_bogus = (int32_t)h->g.nEdgesBogus.size();
}

// Copies the flattened graph into caller arrays (any may be NULL).
DCS_HOST_API void dcs_host_flatten(dcs_host_graph* h, double* pose_xyt, int32_t* edge_a, int32_t* edge_b, double* meas_xyt,
                      uint8_t* kind) {
  ensure_flat(h);
  const FlatGraph& f = h->flat;
  if (pose_xyt) std::memcpy(pose_xyt, f.pose_xyt.data(), f.pose_xyt.size() * sizeof(double));
  if (edge_a) std::memcpy(edge_a, f.edge_a.data(), f.edge_a.size() * sizeof(int32_t));
  if (edge_b) std::memcpy(edge_b, f.edge_b.data(), f.edge_b.size() * sizeof(int32_t));
  if (meas_xyt) std::memcpy(meas_xyt, f.meas_xyt.data(), f.meas_xyt.size() * sizeof(double));
  if (kind) std::memcpy(kind, f.kind.data(), f.kind.size());
}

DCS_HOST_API void dcs_host_set_poses(dcs_host_graph* h, const double* pose_xyt) {
  h->g.scatter_poses(pose_xyt);
  h->flat_valid = false;
}

DCS_HOST_API void dcs_host_write_nodes(dcs_host_graph* h, const char* path) { CoutSilencer s(true); h->g.writePoseGraph_nodes(path); }
DCS_HOST_API void dcs_host_write_edges(dcs_host_graph* h, const char* path) { CoutSilencer s(true); h->g.writePoseGraph_edges(path); }
DCS_HOST_API void dcs_host_write_switches(dcs_host_graph* h, const char* path, const double* priors, const double* optimized, int32_t n) {
  CoutSilencer s(true);
  h->g.writePoseGraph_switches(path, std::vector<double>(priors, priors + n), std::vector<double>(optimized, optimized + n));
}
DCS_HOST_API int dcs_host_write_g2o(dcs_host_graph* h, const char* path) { return synth::write_g2o(h->g, path) ? 0 : 1; }
DCS_HOST_API void dcs_host_graph_free(dcs_host_graph* h) { delete h; }

}  // extern "C"
