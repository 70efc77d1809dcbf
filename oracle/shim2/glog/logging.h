// Stand-in for <glog/logging.h>: the CHECK macros of the reference headers become no-ops that still
// evaluate nothing (test infrastructure; see oracle/Makefile).
#pragma once
#include <iostream>
struct HitlNullStream { template <typename T> HitlNullStream& operator<<(const T&) { return *this; } };
#define CHECK(x) if (false) HitlNullStream()
#define CHECK_EQ(a, b) if (false) HitlNullStream()
#define CHECK_NE(a, b) if (false) HitlNullStream()
#define CHECK_GT(a, b) if (false) HitlNullStream()
#define CHECK_GE(a, b) if (false) HitlNullStream()
#define CHECK_LT(a, b) if (false) HitlNullStream()
#define CHECK_LE(a, b) if (false) HitlNullStream()
#define DCHECK_NE(a, b) if (false) HitlNullStream()
#define DCHECK_EQ(a, b) if (false) HitlNullStream()
#define CHECK_NOTNULL(x) (x)
