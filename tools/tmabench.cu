// Per-SM throughput of TMA tile loads shaped like the head scan's: a (B, 144, HW) bf16 tensor read in tiles of
// all 144 channel rows x W consecutive anchors (one cp.async.bulk.tensor.3d per tile) into a ring of shared-memory
// stages, no consumer work.  Varies W (row length of a request), the ring depth per SM and how many SMs take
// part - the head scan of a step runs on ~84 SMs while the previous step's post stage holds the others.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tmabench tools/tmabench.cu -lcuda
//   tools/tmabench
#include <cuda.h>
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// one CTA per SM (its shared memory sees to that); lane 0 of every warp streams tiles v, v + vgrid, ... (v = its virtual
// CTA number) through a ring of `stages` stages of its own
__global__ void __launch_bounds__(128) tile_read(const __grid_constant__ CUtensorMap map, int tiles_per_stream, int total_tiles,
                                                int tile_w, int tile_bytes, int stages) {
  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ __align__(8) uint64_t bar_all[4][16];
  const int warp = threadIdx.x >> 5, nprod = blockDim.x >> 5;
  unsigned char* smem = smem_all + static_cast<size_t>(warp) * stages * tile_bytes;
  uint64_t* bar = bar_all[warp];
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < stages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int t = blockIdx.x * nprod + warp, issued = 0, waited = 0;
    while (t < total_tiles || waited < issued) {
      if (t < total_tiles && issued - waited < stages) {
        const int s = issued % stages;
        const int b = t / tiles_per_stream, x = (t - b * tiles_per_stream) * tile_w;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(tile_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                smem_u32(smem + static_cast<size_t>(s) * tile_bytes)),
            "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar[s])), "r"(x), "r"(0), "r"(b)
            : "memory");
        ++issued;
        t += gridDim.x * nprod;
      } else {
        const int s = waited % stages;
        const uint32_t parity = (waited / stages) & 1;
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

int main() {
  const int B = 64, CH = 144, HW = 6400 + 1600 + 400;  // the three levels laid out as one 8400-anchor row per channel
  const int nbuf = 8;
  const size_t per = static_cast<size_t>(B) * CH * HW * 2;
  unsigned char* buf;
  CK(cudaMalloc(&buf, per * nbuf));
  CK(cudaMemset(buf, 1, per * nbuf));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(tile_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("TMA tile reads, %.1f MB per launch, %d SMs; columns: tile width (anchors), row bytes, rings per SM x stages, SMs used -> us, GB/s, GB/s per SM\n",
         per / 1e6, sms);
  const int widths[] = {40, 80, 100, 120, 200, 240};
  for (int w : widths) {
    if (HW % w != 0) continue;
    const int tile_bytes = CH * w * 2;
    CUtensorMap maps[nbuf];
    for (int i = 0; i < nbuf; ++i) {
      const cuuint64_t dims[3] = {static_cast<cuuint64_t>(HW), CH, B};
      const cuuint64_t strides[2] = {static_cast<cuuint64_t>(HW) * 2, static_cast<cuuint64_t>(HW) * CH * 2};
      const cuuint32_t box[3] = {static_cast<cuuint32_t>(w), CH, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = cuTensorMapEncodeTiled(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf + per * i, dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        printf("encode failed for width %d (%d)\n", w, static_cast<int>(r));
        return 1;
      }
    }
    const int tps = HW / w, total = tps * B;
    for (int nprod : {1, 3}) {
      int stages = (200 * 1024 / nprod) / tile_bytes;
      if (stages > 12) stages = 12;
      if (stages < 1) continue;
      for (int used : {sms, 84, 64}) {
        const int grid = used;
        const size_t smem = 200 * 1024;  // one CTA per SM, whatever the ring needs
        for (int i = 0; i < 4; ++i) tile_read<<<grid, 32 * nprod, smem>>>(maps[i % nbuf], tps, total, w, tile_bytes, stages);
        CK(cudaDeviceSynchronize());
        float sum = 0.f;
        const int iters = 24;
        for (int i = 0; i < iters; ++i) {
          CK(cudaEventRecord(e0));
          tile_read<<<grid, 32 * nprod, smem>>>(maps[i % nbuf], tps, total, w, tile_bytes, stages);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          sum += ms;
        }
        CK(cudaGetLastError());
        const double us = 1e3 * sum / iters, gbs = per / (us * 1e-6) / 1e9;
        printf("w=%3d row=%3dB  %d rings x %2d stages  SMs=%3d  %7.2f us  %7.1f GB/s  %6.1f GB/s/SM\n", w, w * 2, nprod, stages, used, us,
               gbs, gbs / used);
      }
    }
  }
  return 0;
}
