// Minimal stand-in for <glog/logging.h>, written for this repo (NOT glog code).
// The reference only uses CHECK / CHECK_{EQ,NE,LE,GE} with optional "<< msg"
// streaming; building vendored glog needs cmake + generated headers, so the
// recipe in build_ref.sh puts this directory on the include path instead.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace refshim {
struct Fatal {
  std::ostringstream os;
  Fatal(const char* file, int line, const char* cond) {
    os << file << ":" << line << " CHECK failed: " << cond << " ";
  }
  [[noreturn]] ~Fatal() {
    std::cerr << os.str() << std::endl;
    std::abort();
  }
  template <typename T>
  Fatal& operator<<(const T& v) {
    os << v;
    return *this;
  }
};
struct Voidify {
  void operator&(const Fatal&) {}
};
} // namespace refshim

#define REFSHIM_CHECK(cond, text) \
  (cond) ? (void)0 : ::refshim::Voidify() & ::refshim::Fatal(__FILE__, __LINE__, text)
#define CHECK(c) REFSHIM_CHECK((c), #c)
#define CHECK_EQ(a, b) REFSHIM_CHECK((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) REFSHIM_CHECK((a) != (b), #a " != " #b)
#define CHECK_LE(a, b) REFSHIM_CHECK((a) <= (b), #a " <= " #b)
#define CHECK_LT(a, b) REFSHIM_CHECK((a) < (b), #a " < " #b)
#define CHECK_GE(a, b) REFSHIM_CHECK((a) >= (b), #a " >= " #b)
#define CHECK_GT(a, b) REFSHIM_CHECK((a) > (b), #a " > " #b)
