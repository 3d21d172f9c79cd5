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

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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
    // the file is mapped, not copied: at 1 M poses it is ~0.5 GB of text
    const int fd = ::open(fName.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); return false; }
    const size_t sz = (size_t)st.st_size;
    if (sz == 0) { ::close(fd); return true; }
    void* m = ::mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m != MAP_FAILED) {
      ::madvise(m, sz, MADV_SEQUENTIAL);
      parse(static_cast<const char*>(m), sz);
      ::munmap(m, sz);
      ::close(fd);
      return true;
    }
    ::close(fd);
    FILE* fp = std::fopen(fName.c_str(), "rb");        // not mappable (pipe, odd filesystem): plain read
    if (!fp) return false;
    std::vector<char> buf(sz);
    const size_t got = std::fread(buf.data(), 1, sz, fp);
    std::fclose(fp);
    parse(buf.data(), got);
    return true;
  }

  // Parses g2o text already in memory (also used by the tests).  At 1 M poses the text is ~0.5 GB, so the
  // lines are tokenised and converted by all host threads (chunks cut at line ends, std::from_chars) into
  // plain records, which are then turned into nodes / edges sequentially in file order: same results and
  // the same acceptance rule as a serial reader (an edge is kept iff both ids are below the number of vertex
  // lines read before it).
  void parse(const std::string& text) { parse(text.data(), text.size()); }

  void parse(const char* data, size_t size) {
    struct VRec { int idx; double x, y, th; };
    struct ERec { int a, b; uint32_t vbefore; double v[9]; };
    struct Chunk { std::vector<VRec> v; std::vector<ERec> e; size_t n_near = 0; };
    const char* const end = data + size;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = (unsigned)std::min<size_t>(std::min<unsigned>(nt, 32u), size / (1u << 20) + 1);   // >= 1 MB per thread
    std::vector<const char*> cut(nt + 1, end);
    cut[0] = data;
    for (unsigned c = 1; c < nt; ++c) {
      const char* q = data + size / nt * c;
      if (q < cut[c - 1]) q = cut[c - 1];
      const char* nl = (const char*)std::memchr(q, '\n', (size_t)(end - q));
      cut[c] = nl ? nl + 1 : end;
    }
    std::vector<Chunk> chunks(nt);
    auto work = [&](unsigned c) {
      Chunk& ch = chunks[c];
      const char* s = cut[c];
      const char* const stop = cut[c + 1];
      const char* tok[16];
      uint32_t nv = 0;
      while (s < stop) {
        const char* eol = (const char*)std::memchr(s, '\n', (size_t)(stop - s));
        if (!eol) eol = stop;
        // split on ' ' with compression; a leading space yields an empty first token
        int ntok = 0;
        const char* q = s;
        if (q < eol && *q == ' ') { tok[ntok++] = q; while (q < eol && *q == ' ') ++q; }
        while (q < eol && ntok < 16) {
          tok[ntok++] = q;
          while (q < eol && *q != ' ') ++q;
          while (q < eol && *q == ' ') ++q;
        }
        if (ntok >= 5 && (tag_is(tok[0], eol, "VERTEX_SE2") || tag_is(tok[0], eol, "VERTEX2"))) {
          ch.v.push_back(VRec{to_int(tok[1], eol), to_double(tok[2], eol), to_double(tok[3], eol), to_double(tok[4], eol)});
          ++nv;
        } else if (ntok >= 12 && (tag_is(tok[0], eol, "EDGE_SE2") || tag_is(tok[0], eol, "EDGE2"))) {
          ERec r;
          r.a = to_int(tok[1], eol); r.b = to_int(tok[2], eol); r.vbefore = nv;
          for (int k = 0; k < 9; ++k) r.v[k] = to_double(tok[3 + k], eol);
          ch.n_near += std::abs(r.a - r.b) < 5;
          ch.e.push_back(r);
        }
        s = eol + 1;
      }
    };
    if (nt == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned c = 0; c < nt; ++c) th.emplace_back(work, c);
      for (auto& t : th) t.join();
    }
    size_t tv = 0, te = 0, tn = 0;
    for (const Chunk& ch : chunks) { tv += ch.v.size(); te += ch.e.size(); tn += ch.n_near; }
    nNodes.reserve(nNodes.size() + tv);
    nEdgesOdometry.reserve(nEdgesOdometry.size() + tn);
    nEdgesClosure.reserve(nEdgesClosure.size() + (te - tn));
    for (unsigned c = 0; c < nt; ++c) {
      const Chunk& ch = chunks[c];
      const long long nodes_before = (long long)nNodes.size();
      for (const VRec& r : ch.v) add_node(r.idx, r.x, r.y, r.th);
      for (const ERec& r : ch.e) {
        const long long avail = nodes_before + r.vbefore;     // vertex lines seen before this edge line
        if (r.a < 0 || r.b < 0 || r.a >= avail || r.b >= avail) continue;
        const int type = (std::abs(r.a - r.b) < 5) ? ODOMETRY_EDGE : CLOSURE_EDGE;
        Edge* e = new_edge(nNodes[r.a], nNodes[r.b], type);
        e->setEdgePose(r.v[0], r.v[1], r.v[2]);
        e->setInformationMatrix(r.v[3], r.v[4], r.v[5], r.v[6], r.v[7], r.v[8]);
        (type == ODOMETRY_EDGE ? nEdgesOdometry : nEdgesClosure).push_back(e);
      }
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
    // "index x y theta", 6 significant digits as operator<< / %g print them
    write_parallel(fp, nNodes.size(), 96, [&](size_t i, char* p) {
      const Node* n = nNodes[i];
      p = put_int(p, n->index); *p++ = ' ';
      p = put_g(p, n->p[0]); *p++ = ' ';
      p = put_g(p, n->p[1]); *p++ = ' ';
      p = put_g(p, n->p[2]); *p++ = '\n';
      return p;
    });
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

  // METHOD 2 output (reference :114-148): one line per edge, "a b type prior switch"; odometry edges carry 1 1,
  // loop edges (closure first, then bogus) the prior and the optimised switch value, 6 significant digits.
  void writePoseGraph_switches(const std::string& fname, const std::vector<double>& priors, const std::vector<double>& optimized) {
    std::cout << "#Closure Edges : " << nEdgesClosure.size() << std::endl;
    std::cout << "#Bogus Edges : " << nEdgesBogus.size() << std::endl;
    std::cout << "#priors : " << priors.size() << std::endl;
    std::cout << "#optimized " << optimized.size() << std::endl;
    FILE* fp = std::fopen(fname.c_str(), "w");
    if (!fp) return;
    auto section = [&](const char* title, const std::vector<Edge*>& vec, long offset) {
      std::fputs(title, fp);
      write_parallel(fp, vec.size(), 96, [&](size_t i, char* p) {
        const Edge* e = vec[i];
        const size_t k = (size_t)(offset + (long)i);
        const double pr = offset < 0 ? 1.0 : (k < priors.size() ? priors[k] : 0.0);
        const double sw = offset < 0 ? 1.0 : (k < optimized.size() ? optimized[k] : 0.0);
        p = put_int(p, e->a->index); *p++ = ' ';
        p = put_int(p, e->b->index); *p++ = ' ';
        p = put_int(p, e->edge_type); *p++ = ' ';
        p = put_g(p, pr); *p++ = ' ';
        p = put_g(p, sw); *p++ = '\n';
        return p;
      });
    };
    section("Odometry EDGES AHEAD\n", nEdgesOdometry, -1);
    section("Closure EDGES AHEAD\n", nEdgesClosure, 0);
    section("BOGUS EDGES AHEAD\n", nEdgesBogus, (long)nEdgesClosure.size());
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

  // strtod / strtol semantics (longest valid prefix), through std::from_chars when the token is plain
  // decimal: both are correctly rounded, so the values are identical; anything from_chars does not take
  // whole (a sign '+', hex, a stray suffix) goes to the C function.
  static bool ends_token(const char* p, const char* eol) { return p >= eol || *p == ' ' || *p == '\r' || *p == '\n'; }
  static double to_double(const char* t, const char* eol) {
    double v = 0.0;
    const std::from_chars_result r = std::from_chars(t, eol, v);
    if (r.ec == std::errc() && ends_token(r.ptr, eol)) return v;
    return std::strtod(t, nullptr);
  }
  static int to_int(const char* t, const char* eol) {
    int v = 0;
    const std::from_chars_result r = std::from_chars(t, eol, v);
    if (r.ec == std::errc() && ends_token(r.ptr, eol)) return v;
    return (int)std::strtol(t, nullptr, 10);
  }

  static bool tag_is(const char* t, const char* eol, const char* tag) {
    const size_t n = std::strlen(tag);
    return (size_t)(eol - t) >= n && std::memcmp(t, tag, n) == 0 && (t + n == eol || t[n] == ' ');
  }

  static void write_edges(FILE* fp, const std::vector<Edge*>& vec) {
    write_parallel(fp, vec.size(), 40, [&](size_t i, char* p) {
      const Edge* e = vec[i];
      p = put_int(p, e->a->index); *p++ = ' ';
      p = put_int(p, e->b->index); *p++ = ' ';
      p = put_int(p, e->edge_type); *p++ = '\n';
      return p;
    });
  }

  static char* put_int(char* p, int v) { return std::to_chars(p, p + 12, v).ptr; }
  // printf("%g"): shortest of %e / %f at 6 significant digits, trailing zeros removed
  static char* put_g(char* p, double v) { return std::to_chars(p, p + 24, v, std::chars_format::general, 6).ptr; }

  // Formats n lines with all host threads (each into its own buffer, `max_line` bytes bound per line) and
  // writes the buffers in order.
  template <class F>
  static void write_parallel(FILE* fp, size_t n, size_t max_line, F&& line) {
    if (n == 0) return;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = (unsigned)std::min<size_t>(std::min<unsigned>(nt, 32u), n / 65536 + 1);
    std::vector<std::vector<char>> bufs(nt);
    std::vector<size_t> used(nt, 0);
    auto work = [&](unsigned c) {
      const size_t lo = n * c / nt, hi = n * (c + 1) / nt;
      bufs[c].resize((hi - lo) * max_line);
      char* p = bufs[c].data();
      for (size_t i = lo; i < hi; ++i) p = line(i, p);
      used[c] = (size_t)(p - bufs[c].data());
    };
    if (nt == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned c = 0; c < nt; ++c) th.emplace_back(work, c);
      for (auto& t : th) t.join();
    }
    for (unsigned c = 0; c < nt; ++c) std::fwrite(bufs[c].data(), 1, used[c], fp);
  }
};

#endif
