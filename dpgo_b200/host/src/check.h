// glog-style CHECK: contract violations abort with a message (the reference uses glog CHECK*,
// e.g. src/QuadraticProblem.cpp:30-31), and C-ABI status codes are converted the same way.
#ifndef DPGO_B200_CHECK_H
#define DPGO_B200_CHECK_H
#include <cstdio>
#include <cstdlib>

#include "../../../include/dpgo_b200.h"

#define DPGO_CHECK(cond)                                                                     \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      std::fprintf(stderr, "[DPGO] Check failed: %s at %s:%d\n", #cond, __FILE__, __LINE__); \
      std::abort();                                                                          \
    }                                                                                        \
  } while (0)

#define DPGO_DEVICE_CALL(expr)                                                               \
  do {                                                                                       \
    const int _rc = (expr);                                                                  \
    if (_rc != DPGO_OK) {                                                                    \
      std::fprintf(stderr, "[DPGO] %s failed (%d): %s at %s:%d\n", #expr, _rc,               \
                   dpgo_last_error(), __FILE__, __LINE__);                                   \
      std::abort();                                                                          \
    }                                                                                        \
  } while (0)
#endif
