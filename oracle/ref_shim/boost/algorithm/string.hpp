// Minimal stand-in for <boost/algorithm/string.hpp>: boost::split with is_any_of and
// token_compress_on, as used by the REFERENCE's reader (DCS-ceres/include/g2o_util.h:36).
// TEST INFRASTRUCTURE ONLY — lets oracle/Makefile compile the reference's own g2o_util.h where
// it lies.  Semantics restated from Boost.StringAlgo: runs of separators merge into one; a
// leading / trailing separator yields an empty first / last token; "" yields one empty token.
#ifndef DCS_REF_SHIM_BOOST_STRING
#define DCS_REF_SHIM_BOOST_STRING
#include <string>
namespace boost {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_any_of_pred { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(const std::string& s) { return is_any_of_pred{s}; }
template <typename Seq>
Seq& split(Seq& out, const std::string& in, is_any_of_pred pred, token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  size_t i = 0;
  const size_t n = in.size();
  while (true) {
    while (i < n && !pred(in[i])) cur.push_back(in[i++]);
    out.push_back(cur);
    cur.clear();
    if (i >= n) break;
    ++i;                                                      // the separator
    if (mode == token_compress_on) while (i < n && pred(in[i])) ++i;
  }
  return out;
}
}  // namespace boost
#endif
