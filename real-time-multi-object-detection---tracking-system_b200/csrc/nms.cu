// D1 + N1..N3: YOLOv8 head decode, candidate filter, class-aware NMS, rescale
// (rtm_decode_nms, rtm_nms_pred, rtm_decode_head).
//
// What is reproduced (ultralytics >= 8.1 as called from src/detection/detector.py:100-111;
// SURVEY.md section 3.2 / section 8a rows D1, N1, N2, N3):
//   D1  DFL softmax-expectation over 16 bins x 4 sides, dist2bbox(xywh) * stride, class sigmoid
//   N1  keep anchors whose best class prob > conf; class = FIRST arg-max; then the `classes`
//       filter on that arg-max class; xywh -> xyxy
//   N2  torchvision.ops.nms on boxes + class*7680 (float32 add), stable descending score
//       order (ties: lower candidate index first), suppress when IoU > iou_thres with the
//       float32 IoU compared in double precision, zero-area (NaN IoU) never suppressed,
//       then the first max_det survivors
//   N3  scale_boxes: subtract padding, divide by gain, clip to the source image
//
// Kernel plan
//   decode_candidates  streams the class planes of the three head levels with 16-byte
//                      loads (one thread = 8 bf16/f16 or 4 f32 consecutive anchors, running
//                      max in registers), then the warp decodes the few anchors that can pass
//                      the confidence test co-operatively (144 channels over 32 lanes, DFL
//                      softmax with half-warp shuffles) and appends them to the stream's
//                      candidate list in global memory
//   nms                one CTA per stream: 64-bit key (score desc, anchor asc) bitonic sort
//                      in shared memory, boxes gathered in sorted order, then the greedy scan
//                      run one survivor at a time - every lane tests one later candidate
//                      against the newest survivor and the alive bitmask is rebuilt with warp
//                      ballots (ping-pong buffers, one __syncthreads per survivor) - stopping at
//                      max_det; survivors are rescaled and written in score order.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "rtm_common.cuh"

namespace {

constexpr int kRegMax = 16;
constexpr int kBoxCh = 4 * kRegMax;  // 64
constexpr float kMaxWh = 7680.f;     // ultralytics non_max_suppression max_wh
constexpr int kNmsThreads = 512;
constexpr int kNmsSmemCand = 2048;   // candidates sorted / scanned from shared memory
constexpr int kIdxBits = 15;         // candidate slot (< 32768) in the low key bits
constexpr int kMaxAnchors = 1 << 15;

struct Level {
  int h, w, hw, stride;
  int anchor0;  // first anchor index of the level
};

struct HeadGeom {
  Level lv[3];
  int num_anchors;
  int num_classes;
};

// candidate list of one stream inside the workspace
struct Workspace {
  int32_t* count;     // (B)
  float4* box;        // (B, cap)  xyxy, letterbox pixels
  float* score;       // (B, cap)
  int32_t* meta;      // (B, cap)  anchor | cls << 16
  uint64_t* keys;     // (B, cap_p2)  large-n fallback of the sort
  float4* sbox;       // (B, cap)     large-n fallback of the sorted boxes
  float* sarea;       // (B, cap)
  uint32_t* alive;    // (B, 2, cap/32)
  int cap, cap_p2;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t workspace_layout(int B, int A, char* base, Workspace* ws) {
  const int cap = A, cap_p2 = next_pow2(A), words = (cap + 31) / 32;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_count = take(sizeof(int32_t) * B);
  const size_t o_box = take(sizeof(float4) * B * cap);
  const size_t o_score = take(sizeof(float) * B * cap);
  const size_t o_meta = take(sizeof(int32_t) * B * cap);
  const size_t o_keys = take(sizeof(uint64_t) * B * cap_p2);
  const size_t o_sbox = take(sizeof(float4) * B * cap);
  const size_t o_sarea = take(sizeof(float) * B * cap);
  const size_t o_alive = take(sizeof(uint32_t) * B * 2 * words);
  if (ws) {
    ws->count = reinterpret_cast<int32_t*>(base + o_count);
    ws->box = reinterpret_cast<float4*>(base + o_box);
    ws->score = reinterpret_cast<float*>(base + o_score);
    ws->meta = reinterpret_cast<int32_t*>(base + o_meta);
    ws->keys = reinterpret_cast<uint64_t*>(base + o_keys);
    ws->sbox = reinterpret_cast<float4*>(base + o_sbox);
    ws->sarea = reinterpret_cast<float*>(base + o_sarea);
    ws->alive = reinterpret_cast<uint32_t*>(base + o_alive);
    ws->cap = cap;
    ws->cap_p2 = cap_p2;
  }
  return off;
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ float sigmoidf_rn(float x) {
  return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
}

__device__ __forceinline__ bool class_wanted(const rtm_nms_params& p, int c) {
  return (p.class_mask[c >> 5] >> (c & 31)) & 1u;
}

// dist2bbox(xywh) * stride followed by xywh2xyxy, in the operation order of ultralytics
__device__ __forceinline__ float4 dist_to_xyxy(float l, float t, float r, float b, float ax, float ay,
                                               float stride, float4* xywh) {
  const float x1 = __fsub_rn(ax, l), y1 = __fsub_rn(ay, t);
  const float x2 = __fadd_rn(ax, r), y2 = __fadd_rn(ay, b);
  const float cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.f), stride);
  const float cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.f), stride);
  const float w = __fmul_rn(__fsub_rn(x2, x1), stride);
  const float h = __fmul_rn(__fsub_rn(y2, y1), stride);
  if (xywh) *xywh = make_float4(cx, cy, w, h);
  const float dw = __fdiv_rn(w, 2.f), dh = __fdiv_rn(h, 2.f);
  return make_float4(__fsub_rn(cx, dw), __fsub_rn(cy, dh), __fadd_rn(cx, dw), __fadd_rn(cy, dh));
}

__device__ __forceinline__ void append_candidate(const Workspace& ws, int b, float4 box, float score,
                                                 int cls, int anchor, int32_t* status) {
  const int slot = atomicAdd(&ws.count[b], 1);
  if (slot < ws.cap) {
    const size_t o = static_cast<size_t>(b) * ws.cap + slot;
    ws.box[o] = box;
    ws.score[o] = score;
    ws.meta[o] = anchor | (cls << 16);
  } else if (status) {
    atomicOr(&status[b], RTM_STATUS_CAND_OVERFLOW);
  }
}

// ---------------------------------------------------------------------------------------
// decode_candidates
// ---------------------------------------------------------------------------------------
template <typename T>
struct HeadPtrs {
  const T* p[3];
};

template <typename T, int VEC>
struct alignas(16) Pack {
  T v[VEC];
};

// The warp decodes anchor `pix` of level `lv` of stream b: lane c handles channels
// c, c+32, ... of the 64 + nc channels.  Returns nothing; lane 0 appends the candidate.
template <typename T>
__device__ __forceinline__ void warp_decode_anchor(const T* __restrict__ base, const Level lv, int pix,
                                                   int nc, const rtm_nms_params& prm,
                                                   const Workspace& ws, int b, int32_t* status) {
  const int lane = threadIdx.x & 31;
  // class part: first arg-max of sigmoid over nc classes
  float best = -1.f;
  int bc = 0x7fffffff;
  for (int c = lane; c < nc; c += 32) {
    const float s = sigmoidf_rn(to_float(base[static_cast<size_t>(kBoxCh + c) * lv.hw + pix]));
    if (s > best) {
      best = s;
      bc = c;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float ob = __shfl_xor_sync(rtm::kFull, best, d);
    const int oc = __shfl_xor_sync(rtm::kFull, bc, d);
    if (ob > best || (ob == best && oc < bc)) {
      best = ob;
      bc = oc;
    }
  }
  if (!(best > prm.conf_thres)) return;  // amax(1) > conf_thres, strict
  if (!class_wanted(prm, bc)) return;    // `classes` filter acts on the arg-max class

  // box part: lanes 0-15 / 16-31 hold side l / t (first load) and r / b (second load)
  float d4[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float x = to_float(base[static_cast<size_t>(lane + 32 * k) * lv.hw + pix]);
    float m = x;
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(rtm::kFull, m, d));
    const float e = expf(__fsub_rn(x, m));
    float s = e;
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) s = __fadd_rn(s, __shfl_xor_sync(rtm::kFull, s, d));
    float v = __fmul_rn(static_cast<float>(lane & 15), __fdiv_rn(e, s));
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(rtm::kFull, v, d));
    d4[k] = v;
  }
  const float dl = __shfl_sync(rtm::kFull, d4[0], 0), dt = __shfl_sync(rtm::kFull, d4[0], 16);
  const float dr = __shfl_sync(rtm::kFull, d4[1], 0), db = __shfl_sync(rtm::kFull, d4[1], 16);
  if (lane == 0) {
    const float ax = static_cast<float>(pix % lv.w) + 0.5f, ay = static_cast<float>(pix / lv.w) + 0.5f;
    const float4 box = dist_to_xyxy(dl, dt, dr, db, ax, ay, static_cast<float>(lv.stride), nullptr);
    append_candidate(ws, b, box, best, bc, lv.anchor0 + pix, status);
  }
}

template <typename T, int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS) decode_candidates_kernel(const HeadPtrs<T> heads, const HeadGeom g,
                                                                    const rtm_nms_params prm,
                                                                    const float logit_gate, const Workspace ws,
                                                                    int32_t* status) {
  const int b = blockIdx.y;
  const int grp = blockIdx.x * THREADS + threadIdx.x;  // group of VEC consecutive anchors
  const int a0 = grp * VEC;
  int li = 0;
  if (a0 >= g.lv[1].anchor0) li = 1;
  if (a0 >= g.lv[2].anchor0) li = 2;
  const Level lv = g.lv[li];
  const bool in_range = a0 < g.num_anchors;
  const int pix0 = a0 - lv.anchor0;
  const int nc = g.num_classes;
  const T* base = heads.p[li] + static_cast<size_t>(b) * (kBoxCh + nc) * lv.hw;

  float mx[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) mx[e] = -FLT_MAX;
  if (in_range) {
    const T* cls = base + static_cast<size_t>(kBoxCh) * lv.hw + pix0;
#pragma unroll 8
    for (int c = 0; c < nc; ++c) {
      const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(cls + static_cast<size_t>(c) * lv.hw);
#pragma unroll
      for (int e = 0; e < VEC; ++e) mx[e] = fmaxf(mx[e], to_float(v.v[e]));
    }
  }
  unsigned pending = 0;
#pragma unroll
  for (int e = 0; e < VEC; ++e)
    if (in_range && mx[e] > logit_gate) pending |= 1u << e;

  // warp-cooperative decode of the anchors that passed the gate, in (lane, e) order
  unsigned lanes = __ballot_sync(rtm::kFull, pending != 0);
  while (lanes) {
    const int src = __ffs(lanes) - 1;
    const unsigned bits = __shfl_sync(rtm::kFull, pending, src);
    const int spix0 = __shfl_sync(rtm::kFull, pix0, src);
    const int sli = __shfl_sync(rtm::kFull, li, src);
    const Level slv = g.lv[sli];
    const T* sbase = heads.p[sli] + static_cast<size_t>(b) * (kBoxCh + nc) * slv.hw;
    unsigned rem = bits;
    while (rem) {
      const int e = __ffs(rem) - 1;
      rem &= rem - 1;
      warp_decode_anchor<T>(sbase, slv, spix0 + e, nc, prm, ws, b, status);
    }
    lanes &= lanes - 1;
  }
}

// ---------------------------------------------------------------------------------------
// decode_tma: the production head scan.  Persistent CTAs pull whole tiles (all 64 + nc
// channels x TW consecutive anchors of one stream and level) into shared memory with one
// TMA tensor copy each (cp.async.bulk.tensor.3d, mbarrier completion, kStages deep), so every
// head byte crosses HBM exactly once and no register is tied up by loads in flight.  A tile
// is consumed by 4 x TW threads: thread (anchor a, quarter q) scans a quarter of the class
// rows of its anchor; anchors whose best class passes the confidence test get their four DFL
// sides decoded by their four threads straight from shared memory.
// ---------------------------------------------------------------------------------------
constexpr int kTileW = 80;    // divides 6400 / 1600 / 400 (and every level of a W=640, H%128==0 input)
constexpr int kQuarters = 4;
constexpr int kTmaThreads = kTileW * kQuarters;  // 320

struct TmaGeom {
  HeadGeom g;
  int tiles_before[4];  // tiles of one stream before level l (prefix), [3] = tiles per stream
  int total_tiles;
  int stages;
  int tile_bytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded: a lost TMA completion traps (launch error) instead of hanging the GPU
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int b) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(0), "r"(b)
      : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kTmaThreads) decode_tma_kernel(const __grid_constant__ CUtensorMap map0,
                                                                 const __grid_constant__ CUtensorMap map1,
                                                                 const __grid_constant__ CUtensorMap map2,
                                                                 const TmaGeom tg, const rtm_nms_params prm,
                                                                 const float logit_gate, const Workspace ws,
                                                                 int32_t* status) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ float s_score[kQuarters][kTileW];
  __shared__ int s_class[kQuarters][kTileW];
  __shared__ float s_dist[kQuarters][kTileW];

  const int tid = threadIdx.x;
  const int a = tid % kTileW, q = tid / kTileW;
  const int nc = tg.g.num_classes, stages = tg.stages;
  const int per_q = (nc + kQuarters - 1) / kQuarters;
  const int c_lo = min(q * per_q, nc), c_hi = min(c_lo + per_q, nc);
  const int tps = tg.tiles_before[3];

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int first = blockIdx.x, step = gridDim.x;
  const int my_tiles = first < tg.total_tiles ? (tg.total_tiles - first + step - 1) / step : 0;

  auto issue = [&](int it) {  // thread 0 only
    const int t = first + it * step;
    const int b = t / tps, r = t - b * tps;
    const int li = r >= tg.tiles_before[2] ? 2 : (r >= tg.tiles_before[1] ? 1 : 0);
    const int x = (r - tg.tiles_before[li]) * kTileW;
    const int s = it % stages;
    mbar_expect_tx(&full_bar[s], tg.tile_bytes);
    tma_load_tile(tile_smem + static_cast<size_t>(s) * tg.tile_bytes, li == 0 ? &map0 : (li == 1 ? &map1 : &map2),
                  &full_bar[s], x, b);
  };
  if (tid == 0)
    for (int it = 0; it < min(stages, my_tiles); ++it) issue(it);

  for (int it = 0; it < my_tiles; ++it) {
    const int t = first + it * step;
    const int b = t / tps, r = t - b * tps;
    const int li = r >= tg.tiles_before[2] ? 2 : (r >= tg.tiles_before[1] ? 1 : 0);
    const Level lv = tg.g.lv[li];
    const int pix = (r - tg.tiles_before[li]) * kTileW + a;
    const int s = it % stages;
    mbar_wait(&full_bar[s], (it / stages) & 1);
    const T* tile = reinterpret_cast<const T*>(tile_smem + static_cast<size_t>(s) * tg.tile_bytes);

    // ---- class scan of this thread's quarter: running max, its first index, and the best
    //      value seen BEFORE that index (to detect sigmoid ties between different logits) ----
    float m = -FLT_MAX, m2 = -FLT_MAX;
    int j = c_lo;
#pragma unroll 4
    for (int c = c_lo; c < c_hi; ++c) {
      const float v = to_float(tile[(kBoxCh + c) * kTileW + a]);
      if (v > m) {
        m2 = m;
        m = v;
        j = c;
      }
    }
    float sc = -1.f;
    if (m > logit_gate) {
      sc = sigmoidf_rn(m);
      if (m2 > logit_gate && sigmoidf_rn(m2) == sc) {
        // two different logits round to the same probability: the FIRST class reaching it wins
        for (int c = c_lo; c < j; ++c)
          if (sigmoidf_rn(to_float(tile[(kBoxCh + c) * kTileW + a])) == sc) {
            j = c;
            break;
          }
      }
    }
    s_score[q][a] = sc;
    s_class[q][a] = j;
    __syncthreads();

    // ---- combine the quarters (ascending class order, strict >: first arg-max) ----
    float best = s_score[0][a];
    int bc = s_class[0][a];
#pragma unroll
    for (int k = 1; k < kQuarters; ++k) {
      const float o = s_score[k][a];
      if (o > best) {
        best = o;
        bc = s_class[k][a];
      }
    }
    const bool cand = best > prm.conf_thres && class_wanted(prm, bc);
    if (cand) {
      // DFL side q of this anchor, same operation order as decode_head_kernel
      float x[kRegMax], mx = -FLT_MAX;
#pragma unroll
      for (int k = 0; k < kRegMax; ++k) {
        x[k] = to_float(tile[(q * kRegMax + k) * kTileW + a]);
        mx = fmaxf(mx, x[k]);
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < kRegMax; ++k) {
        x[k] = expf(__fsub_rn(x[k], mx));
        sum = __fadd_rn(sum, x[k]);
      }
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kRegMax; ++k) acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(k), __fdiv_rn(x[k], sum)));
      s_dist[q][a] = acc;
    }
    __syncthreads();  // all reads of the tile are done: its stage can be refilled

    if (tid == 0 && it + stages < my_tiles) issue(it + stages);
    if (cand && q == 0) {
      const float ax = static_cast<float>(pix % lv.w) + 0.5f, ay = static_cast<float>(pix / lv.w) + 0.5f;
      const float4 box = dist_to_xyxy(s_dist[0][a], s_dist[1][a], s_dist[2][a], s_dist[3][a], ax, ay,
                                      static_cast<float>(lv.stride), nullptr);
      append_candidate(ws, b, box, best, bc, lv.anchor0 + pix, status);
    }
  }
}

// ---------------------------------------------------------------------------------------
// decode_head: the full (B, 4 + nc, A) prediction tensor, one thread per anchor.  Used for
// tolerance checks of D1 against the oracle, not on the per-frame path.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void decode_head_kernel(const HeadPtrs<T> heads, const HeadGeom g, float* __restrict__ pred) {
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= g.num_anchors) return;
  int li = 0;
  if (a >= g.lv[1].anchor0) li = 1;
  if (a >= g.lv[2].anchor0) li = 2;
  const Level lv = g.lv[li];
  const int pix = a - lv.anchor0, nc = g.num_classes;
  const T* base = heads.p[li] + static_cast<size_t>(b) * (kBoxCh + nc) * lv.hw + pix;
  float dist[4];
  for (int s = 0; s < 4; ++s) {
    float x[kRegMax], m = -FLT_MAX;
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) {
      x[k] = to_float(base[static_cast<size_t>(s * kRegMax + k) * lv.hw]);
      m = fmaxf(m, x[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) {
      x[k] = expf(__fsub_rn(x[k], m));
      sum = __fadd_rn(sum, x[k]);
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) acc = __fadd_rn(acc, __fmul_rn(static_cast<float>(k), __fdiv_rn(x[k], sum)));
    dist[s] = acc;
  }
  float4 xywh;
  dist_to_xyxy(dist[0], dist[1], dist[2], dist[3], static_cast<float>(pix % lv.w) + 0.5f,
               static_cast<float>(pix / lv.w) + 0.5f, static_cast<float>(lv.stride), &xywh);
  float* out = pred + static_cast<size_t>(b) * (4 + nc) * g.num_anchors + a;
  const size_t A = g.num_anchors;
  out[0] = xywh.x;
  out[A] = xywh.y;
  out[2 * A] = xywh.z;
  out[3 * A] = xywh.w;
  for (int c = 0; c < nc; ++c)
    out[(4 + c) * A] = sigmoidf_rn(to_float(base[static_cast<size_t>(kBoxCh + c) * lv.hw]));
}

// ---------------------------------------------------------------------------------------
// pred_candidates: N1 on an already decoded (B, 4 + nc, A) tensor (rtm_nms_pred)
// ---------------------------------------------------------------------------------------
__global__ void pred_candidates_kernel(const float* __restrict__ pred, int A, int nc, const rtm_nms_params prm,
                                       const Workspace ws, int32_t* status) {
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const float* p = pred + static_cast<size_t>(b) * (4 + nc) * A + a;
  float best = p[4 * static_cast<size_t>(A)];
  int bc = 0;
  for (int c = 1; c < nc; ++c) {
    const float s = p[(4 + c) * static_cast<size_t>(A)];
    if (s > best) {  // strict: first arg-max
      best = s;
      bc = c;
    }
  }
  if (!(best > prm.conf_thres) || !class_wanted(prm, bc)) return;
  const float cx = p[0], cy = p[A], w = p[2 * static_cast<size_t>(A)], h = p[3 * static_cast<size_t>(A)];
  const float dw = __fdiv_rn(w, 2.f), dh = __fdiv_rn(h, 2.f);
  append_candidate(ws, b, make_float4(__fsub_rn(cx, dw), __fsub_rn(cy, dh), __fadd_rn(cx, dw), __fadd_rn(cy, dh)),
                   best, bc, a, status);
}

// ---------------------------------------------------------------------------------------
// nms
// ---------------------------------------------------------------------------------------
struct NmsOut {
  const float* scale;
  float* xyxy;
  float* conf;
  int32_t* cls;
  int32_t* anchor;
  int32_t* keep;
  int32_t* count;
  int32_t stride;
  int32_t* status;
};

__device__ __forceinline__ void bitonic_sort(uint64_t* keys, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += kNmsThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t x = keys[i], y = keys[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            keys[i] = y;
            keys[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kNmsThreads) nms_kernel(const Workspace ws, const rtm_nms_params prm,
                                                          const float iou_gate, const NmsOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_keep[1024];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kNmsThreads / 32;
  int n = ws.count[b];
  if (n > ws.cap) n = ws.cap;  // overflow already flagged by the producer
  const size_t c0 = static_cast<size_t>(b) * ws.cap;
  const int max_det = min(min(prm.max_det, out.stride), 1024);

  if (n <= 0) {
    if (tid == 0) out.count[b] = 0;
    return;
  }
  int P = 1;
  while (P < n) P <<= 1;
  const int words = (n + 31) >> 5;
  const bool in_smem = n <= kNmsSmemCand;
  uint64_t* keys = in_smem ? reinterpret_cast<uint64_t*>(smem_raw) : ws.keys + static_cast<size_t>(b) * ws.cap_p2;
  float4* sbox = in_smem ? reinterpret_cast<float4*>(smem_raw + sizeof(uint64_t) * kNmsSmemCand) : ws.sbox + c0;
  float* sarea = in_smem ? reinterpret_cast<float*>(smem_raw + (sizeof(uint64_t) + sizeof(float4)) * kNmsSmemCand)
                         : ws.sarea + c0;
  uint32_t* alive0 = in_smem ? reinterpret_cast<uint32_t*>(smem_raw + (sizeof(uint64_t) + sizeof(float4) + sizeof(float)) * kNmsSmemCand)
                             : ws.alive + static_cast<size_t>(b) * 2 * ((ws.cap + 31) / 32);
  uint32_t* alive1 = alive0 + (in_smem ? kNmsSmemCand / 32 : (ws.cap + 31) / 32);

  // keys: descending score, then ascending anchor (= torchvision's stable sort of the
  // filtered list, whose order is anchor order); low bits carry the candidate slot
  for (int i = tid; i < P; i += kNmsThreads) {
    uint64_t key = ~0ull;
    if (i < n) {
      const uint32_t s = ~rtm::float_orderable(ws.score[c0 + i]);
      const uint32_t anchor = static_cast<uint32_t>(ws.meta[c0 + i]) & 0xffffu;
      key = (static_cast<uint64_t>(s) << 32) | (static_cast<uint64_t>(anchor) << kIdxBits) | static_cast<uint32_t>(i);
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort(keys, P);

  // gather boxes in sorted order, add the class offset (float32), areas as torchvision does
  for (int i = tid; i < n; i += kNmsThreads) {
    const int slot = static_cast<int>(keys[i] & ((1u << kIdxBits) - 1u));
    float4 bx = ws.box[c0 + slot];
    const int cls = ws.meta[c0 + slot] >> 16;
    const float off = prm.agnostic ? 0.f : __fmul_rn(static_cast<float>(cls), kMaxWh);
    bx.x = __fadd_rn(bx.x, off);
    bx.y = __fadd_rn(bx.y, off);
    bx.z = __fadd_rn(bx.z, off);
    bx.w = __fadd_rn(bx.w, off);
    sbox[i] = bx;
    sarea[i] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
  }
  for (int w = tid; w < words; w += kNmsThreads) {
    const int rem = n - (w << 5);
    alive0[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
  }
  __syncthreads();

  // greedy scan, one survivor per iteration
  uint32_t* cur = alive0;
  uint32_t* nxt = alive1;
  int w0 = 0, kept = 0;
  while (kept < max_det) {
    while (w0 < words && cur[w0] == 0u) ++w0;
    if (w0 >= words) break;
    const int i = (w0 << 5) + __ffs(cur[w0]) - 1;
    const float4 kb = sbox[i];
    const float ka = sarea[i];
    if (tid == 0) s_keep[kept] = i;
    ++kept;
    for (int w = w0 + warp; w < words; w += kWarps) {
      const uint32_t word = cur[w];
      const int j = (w << 5) + lane;
      bool alive = ((word >> lane) & 1u) && j != i;
      if (alive) {
        const float4 jb = sbox[j];
        const float iw = fmaxf(0.f, __fsub_rn(fminf(kb.z, jb.z), fmaxf(kb.x, jb.x)));
        const float ih = fmaxf(0.f, __fsub_rn(fminf(kb.w, jb.w), fmaxf(kb.y, jb.y)));
        const float inter = __fmul_rn(iw, ih);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ka, sarea[j]), inter));
        alive = !(ovr >= iou_gate);  // iou_gate: smallest float32 that is > iou_thres in double
      }
      const uint32_t nw = __ballot_sync(rtm::kFull, alive);
      if (lane == 0) nxt[w] = nw;
    }
    __syncthreads();
    uint32_t* t = cur;
    cur = nxt;
    nxt = t;
  }
  __syncthreads();

  // write the survivors in score order: original box -> scale_boxes -> clip
  float gain = 1.f, padx = 0.f, pady = 0.f, sw = 0.f, sh = 0.f;
  if (out.scale) {
    gain = out.scale[b * 5 + 0];
    padx = out.scale[b * 5 + 1];
    pady = out.scale[b * 5 + 2];
    sw = out.scale[b * 5 + 3];
    sh = out.scale[b * 5 + 4];
  }
  const size_t o0 = static_cast<size_t>(b) * out.stride;
  for (int o = tid; o < kept; o += kNmsThreads) {
    const uint64_t key = keys[s_keep[o]];
    const int slot = static_cast<int>(key & ((1u << kIdxBits) - 1u));
    float4 bx = ws.box[c0 + slot];
    if (out.scale) {
      bx.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.x, padx), gain), 0.f), sw);
      bx.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.y, pady), gain), 0.f), sh);
      bx.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.z, padx), gain), 0.f), sw);
      bx.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.w, pady), gain), 0.f), sh);
    }
    reinterpret_cast<float4*>(out.xyxy)[o0 + o] = bx;
    out.conf[o0 + o] = ws.score[c0 + slot];
    const int meta = ws.meta[c0 + slot];
    out.cls[o0 + o] = meta >> 16;
    if (out.anchor) out.anchor[o0 + o] = meta & 0xffff;
  }
  if (out.keep) {
    // torchvision's index = rank of the survivor's anchor among all candidates
    for (int o = warp; o < kept; o += kWarps) {
      const uint32_t my = static_cast<uint32_t>((keys[s_keep[o]] & 0xffffffffull) >> kIdxBits);
      int cnt = 0;
      for (int i = lane; i < n; i += 32) cnt += (static_cast<uint32_t>(ws.meta[c0 + i]) & 0xffffu) < my;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(rtm::kFull, cnt, d);
      if (lane == 0) out.keep[o0 + o] = cnt;
    }
  }
  if (tid == 0) out.count[b] = kept;
}

constexpr size_t kNmsSmemBytes =
    (sizeof(uint64_t) + sizeof(float4) + sizeof(float)) * kNmsSmemCand + 2 * sizeof(uint32_t) * (kNmsSmemCand / 32);

// smallest float32 g with (double)g > thr, so that `ovr >= g` == `(double)ovr > thr`
// (torchvision compares the float32 IoU with the double threshold)
float iou_gate_for(double thr) {
  const float f = static_cast<float>(thr);
  return static_cast<double>(f) > thr ? f : nextafterf(f, INFINITY);
}

int make_geom(int img_h, int img_w, int nc, HeadGeom* g) {
  RTM_REQUIRE(img_h > 0 && img_w > 0 && img_h % 32 == 0 && img_w % 32 == 0,
              "image size %dx%d must be a positive multiple of 32", img_h, img_w);
  RTM_REQUIRE(nc > 0 && nc <= 256, "num_classes %d out of range (1..256)", nc);
  const int strides[3] = {8, 16, 32};
  int a0 = 0;
  for (int l = 0; l < 3; ++l) {
    Level& lv = g->lv[l];
    lv.stride = strides[l];
    lv.h = img_h / strides[l];
    lv.w = img_w / strides[l];
    lv.hw = lv.h * lv.w;
    lv.anchor0 = a0;
    a0 += lv.hw;
  }
  g->num_anchors = a0;
  g->num_classes = nc;
  RTM_REQUIRE(a0 < kMaxAnchors, "%d anchors exceed the supported %d", a0, kMaxAnchors - 1);
  return RTM_OK;
}

int run_nms(const Workspace& ws, int B, const rtm_nms_params& prm, const NmsOut& out, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    RTM_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kNmsSmemBytes)));
    configured = true;
  }
  {
    rtm::ProfileScope prof(RTM_K_NMS, stream);
    nms_kernel<<<B, kNmsThreads, kNmsSmemBytes, stream>>>(ws, prm, iou_gate_for(prm.iou_thres), out);
  }
  RTM_LAUNCH_CHECK("nms_kernel");
  return RTM_OK;
}

int check_common(const rtm_nms_params* p, int B, float* det_xyxy, float* det_conf, int32_t* det_cls,
                 int32_t* det_count, int det_stride, void* workspace) {
  RTM_REQUIRE(p, "null rtm_nms_params");
  RTM_REQUIRE(B > 0, "num_streams must be positive");
  RTM_REQUIRE(det_xyxy && det_conf && det_cls && det_count && workspace, "null output / workspace pointer");
  RTM_REQUIRE(p->max_det > 0 && p->max_det <= 1024, "max_det %d out of range (1..1024)", p->max_det);
  RTM_REQUIRE(det_stride >= p->max_det, "det_stride %d < max_det %d", det_stride, p->max_det);
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(det_xyxy) & 15) == 0, "det_xyxy must be 16-byte aligned");
  return RTM_OK;
}

float logit_gate_for(float conf_thres) {
  // gate on the raw logit: anything whose sigmoid could exceed conf_thres passes (the exact
  // float32 test is repeated on the sigmoid itself)
  if (conf_thres <= 0.f) return -FLT_MAX;
  if (conf_thres >= 1.f) return FLT_MAX;
  const double lg = log(static_cast<double>(conf_thres) / (1.0 - static_cast<double>(conf_thres)));
  return static_cast<float>(lg - 1e-3 * (1.0 + fabs(lg)));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// which head scan to use: RTM_DECODE_IMPL=tma (default when the shape allows) | ldg
bool want_tma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTM_DECODE_IMPL");
    v = (e && strcmp(e, "ldg") == 0) ? 0 : 1;
  }
  return v == 1;
}

template <typename T>
CUtensorMapDataType tensor_map_dtype();
template <>
CUtensorMapDataType tensor_map_dtype<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
template <>
CUtensorMapDataType tensor_map_dtype<__half>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT16; }
template <>
CUtensorMapDataType tensor_map_dtype<__nv_bfloat16>() { return CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }

// returns 1 when the TMA path was launched, 0 when the caller should fall back, < 0 on error
template <typename T>
int try_launch_decode_tma(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B,
                          const rtm_nms_params& prm, const Workspace& ws, int32_t* status, cudaStream_t stream) {
  if (!want_tma()) return 0;
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) return 0;
  for (int l = 0; l < 3; ++l)
    if (g.lv[l].hw % kTileW != 0) return 0;
  const int ch = kBoxCh + g.num_classes;
  const void* ptrs[3] = {p3, p4, p5};
  CUtensorMap maps[3];
  for (int l = 0; l < 3; ++l) {
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(g.lv[l].hw), static_cast<cuuint64_t>(ch), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(g.lv[l].hw) * sizeof(T),
                                   static_cast<cuuint64_t>(g.lv[l].hw) * ch * sizeof(T)};
    const cuuint32_t box[3] = {kTileW, static_cast<cuuint32_t>(ch), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (ch > 256 || (strides[0] & 15) != 0) return 0;
    const CUresult r = encode(&maps[l], tensor_map_dtype<T>(), 3, const_cast<void*>(ptrs[l]), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 0;
  }
  TmaGeom tg;
  tg.g = g;
  tg.tiles_before[0] = 0;
  for (int l = 0; l < 3; ++l) tg.tiles_before[l + 1] = tg.tiles_before[l] + g.lv[l].hw / kTileW;
  tg.total_tiles = tg.tiles_before[3] * B;
  tg.tile_bytes = ch * kTileW * static_cast<int>(sizeof(T));
  if (tg.tile_bytes % 128 != 0) return 0;
  // 16-bit heads: 4 stages x 22.5 KB, two CTAs per SM; f32 heads: 3 stages x 45 KB, one CTA per SM
  tg.stages = sizeof(T) == 2 ? 4 : 3;
  const int ctas_per_sm = sizeof(T) == 2 ? 2 : 1;
  const size_t smem = static_cast<size_t>(tg.stages) * tg.tile_bytes;
  static size_t configured = 0;
  if (smem > configured) {
    RTM_CUDA(cudaFuncSetAttribute(decode_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  const int grid = min(tg.total_tiles, rtm::sm_count() * ctas_per_sm);
  {
    rtm::ProfileScope prof(RTM_K_DECODE, stream);
    decode_tma_kernel<T><<<grid, kTmaThreads, smem, stream>>>(maps[0], maps[1], maps[2], tg, prm,
                                                             logit_gate_for(prm.conf_thres), ws, status);
  }
  RTM_LAUNCH_CHECK("decode_tma_kernel");
  return 1;
}

template <typename T, int VEC>
int launch_decode(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B,
                  const rtm_nms_params& prm, const Workspace& ws, int32_t* status, cudaStream_t stream) {
  for (int l = 0; l < 3; ++l)
    RTM_REQUIRE(g.lv[l].hw % VEC == 0, "level %d has %d anchors, not a multiple of %d", l, g.lv[l].hw, VEC);
  HeadPtrs<T> heads{{static_cast<const T*>(p3), static_cast<const T*>(p4), static_cast<const T*>(p5)}};
  for (int l = 0; l < 3; ++l)
    RTM_REQUIRE((reinterpret_cast<uintptr_t>(heads.p[l]) & 15) == 0, "head level %d must be 16-byte aligned", l);
  const int tma = try_launch_decode_tma<T>(p3, p4, p5, g, B, prm, ws, status, stream);
  if (tma != 0) return tma < 0 ? tma : RTM_OK;
  const float gate = logit_gate_for(prm.conf_thres);
  constexpr int THREADS = 128;
  const int groups = g.num_anchors / VEC;
  dim3 grid((groups + THREADS - 1) / THREADS, B);
  {
    rtm::ProfileScope prof(RTM_K_DECODE, stream);
    decode_candidates_kernel<T, VEC, THREADS><<<grid, THREADS, 0, stream>>>(heads, g, prm, gate, ws, status);
  }
  RTM_LAUNCH_CHECK("decode_candidates_kernel");
  return RTM_OK;
}

}  // namespace

extern "C" size_t rtm_nms_workspace_bytes(int32_t num_streams, int32_t num_anchors) {
  if (num_streams <= 0 || num_anchors <= 0) return 0;
  return workspace_layout(num_streams, num_anchors, nullptr, nullptr);
}

extern "C" int rtm_decode_nms(const void* head_p3, const void* head_p4, const void* head_p5,
                              int32_t head_dtype, int32_t num_streams, int32_t img_h, int32_t img_w,
                              const rtm_nms_params* params, const float* scale, float* det_xyxy,
                              float* det_conf, int32_t* det_cls, int32_t* det_anchor, int32_t* det_keep,
                              int32_t* det_count, int32_t det_stride, int32_t* status, void* workspace,
                              size_t workspace_bytes, rtm_cuda_stream stream) {
  int rc = check_common(params, num_streams, det_xyxy, det_conf, det_cls, det_count, det_stride, workspace);
  if (rc) return rc;
  RTM_REQUIRE(head_p3 && head_p4 && head_p5, "rtm_decode_nms: null head tensor");
  HeadGeom g;
  rc = make_geom(img_h, img_w, params->num_classes, &g);
  if (rc) return rc;
  Workspace ws;
  const size_t need = workspace_layout(num_streams, g.num_anchors, static_cast<char*>(workspace), &ws);
  RTM_REQUIRE(workspace_bytes >= need, "rtm_decode_nms: workspace has %zu bytes, %zu needed", workspace_bytes, need);
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RTM_CUDA(cudaMemsetAsync(ws.count, 0, sizeof(int32_t) * num_streams, s));
  switch (head_dtype) {
    case RTM_F32:
      rc = launch_decode<float, 4>(head_p3, head_p4, head_p5, g, num_streams, *params, ws, status, s);
      break;
    case RTM_F16:
      rc = launch_decode<__half, 8>(head_p3, head_p4, head_p5, g, num_streams, *params, ws, status, s);
      break;
    case RTM_BF16:
      rc = launch_decode<__nv_bfloat16, 8>(head_p3, head_p4, head_p5, g, num_streams, *params, ws, status, s);
      break;
    default:
      RTM_REQUIRE(false, "rtm_decode_nms: unknown head_dtype %d", head_dtype);
  }
  if (rc) return rc;
  NmsOut out{scale, det_xyxy, det_conf, det_cls, det_anchor, det_keep, det_count, det_stride, status};
  return run_nms(ws, num_streams, *params, out, s);
}

extern "C" int rtm_nms_pred(const float* pred, int32_t num_streams, int32_t num_anchors,
                            const rtm_nms_params* params, const float* scale, float* det_xyxy,
                            float* det_conf, int32_t* det_cls, int32_t* det_anchor, int32_t* det_keep,
                            int32_t* det_count, int32_t det_stride, int32_t* status, void* workspace,
                            size_t workspace_bytes, rtm_cuda_stream stream) {
  int rc = check_common(params, num_streams, det_xyxy, det_conf, det_cls, det_count, det_stride, workspace);
  if (rc) return rc;
  RTM_REQUIRE(pred, "rtm_nms_pred: null prediction tensor");
  RTM_REQUIRE(num_anchors > 0 && num_anchors < kMaxAnchors, "rtm_nms_pred: num_anchors %d out of range", num_anchors);
  RTM_REQUIRE(params->num_classes > 0 && params->num_classes <= 256, "num_classes out of range");
  Workspace ws;
  const size_t need = workspace_layout(num_streams, num_anchors, static_cast<char*>(workspace), &ws);
  RTM_REQUIRE(workspace_bytes >= need, "rtm_nms_pred: workspace has %zu bytes, %zu needed", workspace_bytes, need);
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RTM_CUDA(cudaMemsetAsync(ws.count, 0, sizeof(int32_t) * num_streams, s));
  dim3 grid((num_anchors + 255) / 256, num_streams);
  {
    rtm::ProfileScope prof(RTM_K_PRED, s);
    pred_candidates_kernel<<<grid, 256, 0, s>>>(pred, num_anchors, params->num_classes, *params, ws, status);
  }
  RTM_LAUNCH_CHECK("pred_candidates_kernel");
  NmsOut out{scale, det_xyxy, det_conf, det_cls, det_anchor, det_keep, det_count, det_stride, status};
  return run_nms(ws, num_streams, *params, out, s);
}

extern "C" int rtm_decode_head(const void* head_p3, const void* head_p4, const void* head_p5,
                               int32_t head_dtype, int32_t num_streams, int32_t img_h, int32_t img_w,
                               int32_t num_classes, float* pred, rtm_cuda_stream stream) {
  RTM_REQUIRE(head_p3 && head_p4 && head_p5 && pred, "rtm_decode_head: null pointer");
  RTM_REQUIRE(num_streams > 0, "num_streams must be positive");
  HeadGeom g;
  int rc = make_geom(img_h, img_w, num_classes, &g);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid((g.num_anchors + 127) / 128, num_streams);
  switch (head_dtype) {
    case RTM_F32: {
      HeadPtrs<float> h{{static_cast<const float*>(head_p3), static_cast<const float*>(head_p4), static_cast<const float*>(head_p5)}};
      decode_head_kernel<float><<<grid, 128, 0, s>>>(h, g, pred);
      break;
    }
    case RTM_F16: {
      HeadPtrs<__half> h{{static_cast<const __half*>(head_p3), static_cast<const __half*>(head_p4), static_cast<const __half*>(head_p5)}};
      decode_head_kernel<__half><<<grid, 128, 0, s>>>(h, g, pred);
      break;
    }
    case RTM_BF16: {
      HeadPtrs<__nv_bfloat16> h{{static_cast<const __nv_bfloat16*>(head_p3), static_cast<const __nv_bfloat16*>(head_p4), static_cast<const __nv_bfloat16*>(head_p5)}};
      decode_head_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(h, g, pred);
      break;
    }
    default:
      RTM_REQUIRE(false, "rtm_decode_head: unknown head_dtype %d", head_dtype);
  }
  RTM_LAUNCH_CHECK("decode_head_kernel");
  return RTM_OK;
}
