// shim (test support only): the CHECK family the reference's tests use, aborting like glog does.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iostream>

namespace shim_glog {
struct Voidify {
  void operator&(std::ostream&) {}
};
struct Fatal {
  Fatal(const char* file, int line, const char* what) { std::cerr << "CHECK failed: " << what << " at " << file << ":" << line << " "; }
  [[noreturn]] ~Fatal() {
    std::cerr << std::endl;
    std::abort();
  }
  std::ostream& stream() { return std::cerr; }
};
}  // namespace shim_glog
#define CHECK(c) (c) ? (void)0 : shim_glog::Voidify() & shim_glog::Fatal(__FILE__, __LINE__, #c).stream()
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
