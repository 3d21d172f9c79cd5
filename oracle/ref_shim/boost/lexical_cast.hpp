// Minimal stand-in for <boost/lexical_cast.hpp> (strict whole-string conversion, throws on junk),
// as used by the REFERENCE's reader (DCS-ceres/include/g2o_util.h:40-66).  TEST INFRASTRUCTURE ONLY.
#ifndef DCS_REF_SHIM_BOOST_LEXICAL_CAST
#define DCS_REF_SHIM_BOOST_LEXICAL_CAST
#include <cstdlib>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <typename T> T lexical_cast(const std::string& s);
template <> inline int lexical_cast<int>(const std::string& s) {
  char* end = nullptr;
  const long v = std::strtol(s.c_str(), &end, 10);
  if (s.empty() || *end != '\0') throw bad_lexical_cast();
  return (int)v;
}
template <> inline double lexical_cast<double>(const std::string& s) {
  char* end = nullptr;
  const double v = std::strtod(s.c_str(), &end);
  if (s.empty() || *end != '\0') throw bad_lexical_cast();
  return v;
}
}  // namespace boost
#endif
