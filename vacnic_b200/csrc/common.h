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

// Programmatic dependent launch (PDL).  Kernels launched through launch_pdl carry
// cudaLaunchAttributeProgrammaticStreamSerialization: the grid may be scheduled while its predecessor in the stream is
// still draining, runs its prologue (barrier init, TMEM allocation, index math) and then blocks in pdl_wait() -- which
// every such kernel executes BEFORE its first global-memory access -- until the predecessor has completed and flushed.
// pdl_trigger() follows pdl_wait() in every kernel, so a grid overlaps with its direct predecessor only.  Inside a
// captured CUDA graph the same launches become programmatic dependency edges.  VACNIC_PDL=0 turns the attribute off
// (pdl_wait / pdl_trigger are then no-ops).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through check_last()
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {  // first thing before any global-memory access
  pdl_wait();
  pdl_trigger();
}
#endif

#define VB_REQUIRE(cond, ...)                       \
  do {                                              \
    if (!(cond)) return vb::fail(VACNIC_EINVAL, __VA_ARGS__); \
  } while (0)

}  // namespace vb
