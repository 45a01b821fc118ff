// Host-side helpers shared by the C-ABI translation units: error reporting, launch accounting,
// device properties.  No torch types anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vacnic_b200.h"

namespace vb {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();          // SMs of the current device (cached per device)
int check_last(const char* what);  // cudaGetLastError -> VACNIC_ECUDA

#define VB_REQUIRE(cond, ...)                       \
  do {                                              \
    if (!(cond)) return vb::fail(VACNIC_EINVAL, __VA_ARGS__); \
  } while (0)

}  // namespace vb
