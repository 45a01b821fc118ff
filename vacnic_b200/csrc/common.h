// Host-side helpers shared by the C-ABI translation units: error reporting, launch accounting,
// device properties.  No torch types anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vacnic_b200.h"

struct CUtensorMap_st;  // <cuda.h>

namespace vb {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();          // SMs of the current device (cached per device)
int check_last(const char* what);  // cudaGetLastError -> VACNIC_ECUDA

// 4-D bf16 tensor map (SWIZZLE_128B) for one MMA operand, built in gemm_sm100.cu.  K-major: dims
// (K, rows, b0, b1), box (64, box_rows).  MN-major: dims (rows, K, b0, b1), box (64, 64).
int make_operand_map(::CUtensorMap_st* tm, const void* base, bool mn_major, int rows, int K, long long ld, int batch0,
                     long long sb0, int batch1, long long sb1, int box_rows);

#define VB_REQUIRE(cond, ...)                       \
  do {                                              \
    if (!(cond)) return vb::fail(VACNIC_EINVAL, __VA_ARGS__); \
  } while (0)

}  // namespace vb
