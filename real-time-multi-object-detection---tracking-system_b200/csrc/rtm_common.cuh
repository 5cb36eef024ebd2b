// Shared helpers for the sm_100a kernels of librtmodt_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "rtmodt_b200.h"

namespace rtm {

// ---- error plumbing (host) ------------------------------------------------------------
void set_error(const char* fmt, ...);

#define RTM_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::rtm::set_error(__VA_ARGS__);    \
      return RTM_ERR_INVALID;           \
    }                                   \
  } while (0)

#define RTM_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t err__ = (call);                                                   \
    if (err__ != cudaSuccess) {                                                   \
      ::rtm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                       __FILE__, __LINE__);                                       \
      return RTM_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define RTM_LAUNCH_CHECK(name)                                                         \
  do {                                                                                 \
    cudaError_t err__ = cudaGetLastError();                                            \
    if (err__ != cudaSuccess) {                                                        \
      ::rtm::set_error("launch of %s failed: %s", name, cudaGetErrorString(err__));    \
      return RTM_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

int sm_count();  // cached multiProcessorCount of the current device
// raises a kernel's dynamic shared-memory limit when `bytes` exceeds what was set for it on the current device
int ensure_dynamic_smem(const void* func, size_t bytes);
bool pdl_enabled();  // programmatic dependent launch of the step kernels (RTM_PDL=0 turns it off)

// Brackets a kernel launch with CUDA events while rtm_profile_enable(1) is in effect.
extern bool g_profile_on;
void profile_begin(int kind, cudaStream_t s);
void profile_end(cudaStream_t s);
struct ProfileScope {
  cudaStream_t s;
  bool on;
  ProfileScope(int kind, cudaStream_t stream) : s(stream), on(g_profile_on) {
    if (on) profile_begin(kind, s);
  }
  ~ProfileScope() {
    if (on) profile_end(s);
  }
};

// ---- device helpers -------------------------------------------------------------------
// Diagnosis builds (-DRTM_TIMELINE, tools/post_timeline.py): thread 0 of every CTA stamps
// %globaltimer at the stage boundaries of the post kernel.  Compiled out of the product library.
#ifdef RTM_TIMELINE
static __device__ unsigned long long* g_timeline = nullptr;
__device__ __forceinline__ void timeline_mark(int i) {
  if (threadIdx.x == 0 && g_timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_timeline[blockIdx.x * 32 + i] = t;
  }
}
#define RTM_TL(i) ::rtm::timeline_mark(i)
#else
#define RTM_TL(i) ((void)0)
#endif

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Order-preserving compaction support: exclusive prefix of `flag` over the threads of the
// block (thread order), plus the block total.  `scratch` holds 33 ints.
// The block's first THREADS threads - all that are still running - must call it; contains three __syncthreads().
template <int THREADS>
__device__ __forceinline__ int block_exclusive_count(bool flag, int* scratch, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarp = (THREADS + 31) >> 5;  // the threads that take part (a block may hold more: they have left by then)
  const unsigned bal = __ballot_sync(kFull, flag);
  const int within = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) scratch[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    // nwarp <= 32: one warp scans the per-warp counts
    int v = lane < nwarp ? scratch[lane] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane < nwarp) scratch[lane] = incl - v;
    if (lane == 31) scratch[32] = incl;
  }
  __syncthreads();
  const int base = scratch[warp];
  *total = scratch[32];
  __syncthreads();  // scratch may be reused by the next call
  return base + within;
}

// Exclusive prefix sum of `v` over the threads of the block (thread order) plus the block
// total.  `scratch` holds 33 ints.  All threads must call it; three __syncthreads().
template <int THREADS>
__device__ __forceinline__ int block_exclusive_sum(int v, int* scratch, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarp = (THREADS + 31) >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < nwarp ? scratch[lane] : 0;
    int inc2 = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(kFull, inc2, d);
      if (lane >= d) inc2 += o;
    }
    if (lane < nwarp) scratch[lane] = inc2 - w;
    if (lane == 31) scratch[32] = inc2;
  }
  __syncthreads();
  const int base = scratch[warp];
  *total = scratch[32];
  __syncthreads();
  return base + incl - v;
}

// total order on floats as unsigned ints (ascending)
__device__ __forceinline__ uint32_t float_orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_orderable(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace rtm
