// Read-only bandwidth probes for the size of one decode launch (154.8 MB): what a launch of
// this size can reach on a B200 with (a) plain coalesced 16-byte loads, (b) 1-D bulk copies
// global -> shared (UBLKCP) and nothing else.  Gives the practical ceiling the head scan is
// compared with beside the long-copy figure of MEASURED_PEAKS.json.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu
//   tools/membench [MB per launch = 154.8288] [buffers = 16]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                        \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__global__ void __launch_bounds__(256) ldg_sum(const uint4* __restrict__ p, size_t n16, unsigned* out) {
  unsigned acc = 0;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
    acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
  }
  for (; i < n16; i += stride) {
    const uint4 a = __ldcs(p + i);
    acc += a.x ^ a.y ^ a.z ^ a.w;
  }
  if (acc == 0x12345678u) *out = acc;  // never true in practice; keeps the loads alive
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// persistent CTAs; one thread streams CHUNK-byte pieces into a ring of shared-memory stages with
// cp.async.bulk and waits for each; no consumer work at all
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(32) bulk_read(const unsigned char* __restrict__ p, size_t chunks) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    size_t c = blockIdx.x;
    int issued = 0, waited = 0;
    while (c < chunks || waited < issued) {
      if (c < chunks && issued - waited < STAGES) {
        const int s = issued % STAGES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + static_cast<size_t>(s) * CHUNK)),
                     "l"(p + c * CHUNK), "r"(CHUNK), "r"(smem_u32(&bar[s]))
                     : "memory");
        ++issued;
        c += gridDim.x;
      } else {
        const int s = waited % STAGES;
        const uint32_t parity = (waited / STAGES) & 1;
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                       : "=r"(ok)
                       : "r"(smem_u32(&bar[s])), "r"(parity)
                       : "memory");
        ++waited;
      }
    }
  }
}

int main(int argc, char** argv) {
  const double mb = argc > 1 ? atof(argv[1]) : 154.8288;
  const int nbuf = argc > 2 ? atoi(argv[2]) : 16;
  const size_t bytes = static_cast<size_t>(mb * 1e6) / 65536 * 65536;
  unsigned char* buf;
  unsigned* out;
  CK(cudaMalloc(&buf, bytes * nbuf));
  CK(cudaMalloc(&out, 4));
  CK(cudaMemset(buf, 1, bytes * nbuf));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto time_it = [&](const char* name, auto launch) {
    for (int i = 0; i < 8; ++i) launch(i % nbuf);
    CK(cudaDeviceSynchronize());
    const int iters = 64;
    float best = 1e9f, sum = 0.f;
    for (int i = 0; i < iters; ++i) {
      CK(cudaEventRecord(e0));
      launch(i % nbuf);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best;
      sum += ms;
    }
    CK(cudaGetLastError());
    printf("%-28s avg %7.2f us  %7.1f GB/s   best %7.2f us  %7.1f GB/s\n", name, 1e3 * sum / iters,
           bytes / (sum / iters * 1e-3) / 1e9, 1e3 * best, bytes / (best * 1e-3) / 1e9);
  };
  printf("read-only probes, %.1f MB per launch, %d distinct buffers (%.1f GB), %d SMs\n", bytes / 1e6, nbuf,
         bytes * nbuf / 1e9, sms);
  for (int per_sm : {4, 8, 16}) {
    char name[64];
    snprintf(name, sizeof name, "ldg.128 x4, %d CTAs/SM", per_sm);
    time_it(name, [&](int b) { ldg_sum<<<sms * per_sm, 256>>>(reinterpret_cast<const uint4*>(buf + bytes * b), bytes / 16, out); });
  }
  {
    CK(cudaFuncSetAttribute(bulk_read<16384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
    CK(cudaFuncSetAttribute(bulk_read<32768, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 3));
    CK(cudaFuncSetAttribute(bulk_read<8192, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
    time_it("bulk 16K x4 stages, 3/SM", [&](int b) { bulk_read<16384, 4><<<sms * 3, 32, 16384 * 4>>>(buf + bytes * b, bytes / 16384); });
    time_it("bulk 32K x3 stages, 2/SM", [&](int b) { bulk_read<32768, 3><<<sms * 2, 32, 32768 * 3>>>(buf + bytes * b, bytes / 32768); });
    time_it("bulk 8K x8 stages, 3/SM", [&](int b) { bulk_read<8192, 8><<<sms * 3, 32, 8192 * 8>>>(buf + bytes * b, bytes / 8192); });
  }
  // the long-copy reference point (what MEASURED_PEAKS.json times): one 1 GiB read
  if (bytes * nbuf >= (1ull << 30)) {
    const size_t big = 1ull << 30;
    for (int i = 0; i < 3; ++i) ldg_sum<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(buf), big / 16, out);
    CK(cudaEventRecord(e0));
    ldg_sum<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(buf), big / 16, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-28s %7.2f us  %7.1f GB/s\n", "ldg.128, 1 GiB read", 1e3 * ms, big / (ms * 1e-3) / 1e9);
  }
  return 0;
}
