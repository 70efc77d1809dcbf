// Stand-in for <glog/logging.h> (shim3 build of the reference's .cpp files): CHECK macros that abort with a message,
// as glog's do, so a violated reference precondition is visible to the parity tests.  Test infrastructure.
#pragma once
#include <cstdlib>
#include <iostream>
struct HitlFatalStream {
  bool live;
  explicit HitlFatalStream(bool l) : live(l) {}
  ~HitlFatalStream() { if (live) { std::cerr << std::endl; std::abort(); } }
  template <typename T> HitlFatalStream& operator<<(const T& v) { if (live) std::cerr << v; return *this; }
};
#define HITL_CHECK_OP(a, op, b) if ((a) op (b)) {} else HitlFatalStream(true) << "CHECK failed: " #a " " #op " " #b " "
#define CHECK(x) if (x) {} else HitlFatalStream(true) << "CHECK failed: " #x " "
#define CHECK_EQ(a, b) HITL_CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) HITL_CHECK_OP(a, !=, b)
#define CHECK_GT(a, b) HITL_CHECK_OP(a, >, b)
#define CHECK_GE(a, b) HITL_CHECK_OP(a, >=, b)
#define CHECK_LT(a, b) HITL_CHECK_OP(a, <, b)
#define CHECK_LE(a, b) HITL_CHECK_OP(a, <=, b)
#define DCHECK(x) if (true) {} else HitlFatalStream(false)
#define DCHECK_NE(a, b) if (true) {} else HitlFatalStream(false)
#define DCHECK_EQ(a, b) if (true) {} else HitlFatalStream(false)
#define DCHECK_GT(a, b) if (true) {} else HitlFatalStream(false)
#define DCHECK_LT(a, b) if (true) {} else HitlFatalStream(false)
#define CHECK_NOTNULL(x) (x)
#define LOG(x) HitlFatalStream(false)
