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
// Kernel plan (numbers and experiments: DESIGN.md section 5)
//   decode_tma   the production head scan.  Persistent CTAs; a producer warp pulls whole tiles
//                (all 64 + nc channels x 80 consecutive anchors of one stream and level) into a
//                ring of shared-memory stages with one TMA tensor copy each (full / empty
//                mbarriers, L2 evict-first), so every head byte crosses HBM exactly once.  Tiles
//                are handed out by a ticket counter after a static first ring round, so that any
//                set of resident CTAs shares the work - the scan runs beside the previous step's
//                post kernel (programmatic dependent launch) or on a stream of its own
//                (scan_async).  Five consumer warps work independently of each other (no block
//                barrier): a warp owns 16 anchors of the tile, a lane owns two adjacent anchors x
//                one class quarter and keeps a packed (bf16x2 / f16x2) running maximum over its
//                class rows, bank-conflict free; the four lanes of an anchor combine by shuffle,
//                and anchors that pass the confidence test get their four DFL sides decoded by
//                their four lanes from shared memory.  SPLIT variant (RTM_TMA_SPLIT=1): class
//                planes in the ring, box planes as per-warp sub-tiles fetched only for candidates.
//   decode_scan  no ring: streams the class planes with 16-byte loads, reads the box planes only
//                for the 8-anchor groups that hold a candidate; the path for shapes the tiling
//                does not cover (RTM_DECODE_IMPL=scan forces it).
//   decode_ldg   the first version of the scan, kept for A/B runs (RTM_DECODE_IMPL=ldg).
//   candidates   are written dense by anchor plus a bitmask into one slot of a ring of three (see
//                nms_body.cuh): no atomics, no per-frame reset, and the list order is ultralytics'
//                filtered-tensor order.
//   nms          one CTA per stream (nms_body.cuh): counting-rank sort of (class, score desc, rank
//                asc) keys, class-parallel greedy scan (warp heads, parallel filter, warp tails;
//                block-wide scan for long segments, agnostic NMS and boxes outside the class
//                guard), first max_det survivors by score, rescale, write in score order.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "nms_body.cuh"

namespace {

using rtm::kFull;
using rtm::NmsOut;
using rtm::Workspace;

constexpr int kRegMax = 16;
constexpr int kBoxCh = 4 * kRegMax;  // 64
constexpr int kNmsThreads = 512;

struct Level {
  int h, w, hw, stride;
  int anchor0;  // first anchor index of the level
};

struct HeadGeom {
  Level lv[3];
  int num_anchors;
  int num_classes;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Workspace = header (one tile ticket counter per slot) + a ring of kCandSlots copies of the
// candidate interchange arrays + one set of spill arrays.  Consecutive head scans take
// consecutive slots, so the scan of step k+2 may already run while the post kernel of step k
// still reads its own slot (programmatic dependent launch: see decode_tma_kernel and post.cu).
using rtm::kCandSlots;
constexpr size_t kWorkspaceHeader = 128 * kCandSlots;

size_t workspace_layout(int B, int A, char* base, Workspace* ws, int half = 0) {
  const int words = (A + 31) / 32, cap_p2 = next_pow2(A);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_head = take(kWorkspaceHeader);
  size_t o_mask[kCandSlots], o_box[kCandSlots], o_score[kCandSlots], o_cls[kCandSlots];
  for (int h = 0; h < kCandSlots; ++h) {
    o_mask[h] = take(sizeof(uint32_t) * B * words);
    o_box[h] = take(sizeof(float4) * B * A);
    o_score[h] = take(sizeof(float) * B * A);
    o_cls[h] = take(sizeof(int32_t) * B * A);
  }
  const size_t o_keys = take(sizeof(uint64_t) * B * cap_p2);
  const size_t o_ubox = take(sizeof(float4) * B * A);
  const size_t o_sbox = take(sizeof(float4) * B * A);
  const size_t o_sarea = take(sizeof(float) * B * A);
  const size_t o_loc = take(sizeof(int32_t) * B * A);
  const size_t o_alive = take(sizeof(uint32_t) * B * 2 * words);
  if (ws) {
    ws->tile_counter = reinterpret_cast<int*>(base + o_head) + 32 * half;  // 128 bytes apart
    ws->mask = reinterpret_cast<uint32_t*>(base + o_mask[half]);
    ws->box = reinterpret_cast<float4*>(base + o_box[half]);
    ws->score = reinterpret_cast<float*>(base + o_score[half]);
    ws->cls = reinterpret_cast<int32_t*>(base + o_cls[half]);
    ws->keys = reinterpret_cast<uint64_t*>(base + o_keys);
    ws->ubox = reinterpret_cast<float4*>(base + o_ubox);
    ws->sbox = reinterpret_cast<float4*>(base + o_sbox);
    ws->sarea = reinterpret_cast<float*>(base + o_sarea);
    ws->loc = reinterpret_cast<int32_t*>(base + o_loc);
    ws->alive = reinterpret_cast<uint32_t*>(base + o_alive);
    ws->num_anchors = A;
    ws->words = words;
    ws->cap_p2 = cap_p2;
    ws->slot = half;
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
  // x / 2 == x * 0.5f bit for bit (exact scaling by a power of two); the division sequence is ~10x the instructions
  const float cx = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), stride);
  const float cy = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), stride);
  const float w = __fmul_rn(__fsub_rn(x2, x1), stride);
  const float h = __fmul_rn(__fsub_rn(y2, y1), stride);
  if (xywh) *xywh = make_float4(cx, cy, w, h);
  const float dw = __fmul_rn(w, 0.5f), dh = __fmul_rn(h, 0.5f);
  return make_float4(__fsub_rn(cx, dw), __fsub_rn(cy, dh), __fadd_rn(cx, dw), __fadd_rn(cy, dh));
}

// DFL expectation of one side from 16 logits: sum_k k * softmax(x)_k.  Shared by all decode
// kernels so that they agree bit for bit; checked against the oracle within 1e-4 relative
// (D1 is the tolerance-checked stage: torch's CPU softmax rounds differently anyway), hence
// the fast exponential and a single division.
__device__ __forceinline__ float dfl_expectation(float (&x)[kRegMax]) {
  float mx = x[0];
#pragma unroll
  for (int k = 1; k < kRegMax; ++k) mx = fmaxf(mx, x[k]);
  float sum = 0.f, acc = 0.f;
#pragma unroll
  for (int k = 0; k < kRegMax; ++k) {
    const float e = __expf(x[k] - mx);
    sum += e;
    acc = __fmaf_rn(static_cast<float>(k), e, acc);
  }
  return __fdiv_rn(acc, sum);
}

__device__ __forceinline__ void store_candidate(const Workspace& ws, int b, int anchor, float4 box, float score, int cls) {
  const size_t o = static_cast<size_t>(b) * ws.num_anchors + anchor;
  ws.box[o] = box;
  ws.score[o] = score;
  ws.cls[o] = cls;
}

template <typename T>
struct HeadPtrs {
  const T* p[3];
};

// ---------------------------------------------------------------------------------------
// decode_tma
// ---------------------------------------------------------------------------------------
// Tile width (anchors per tile) is a template parameter: 64 / 128 make every channel row of a
// 16-bit tile a whole number of 128-byte lines (one or two full-line L2 requests per row), 80
// divides 6400 / 1600 / 400 exactly but gives 160-byte rows that straddle lines.  Tiles that run
// past the end of a level are zero-filled by the TMA unit and their anchors masked out.
constexpr int kAnchorsPerWarp = 16;  // a lane owns 2 adjacent anchors x one class quarter
constexpr int kMaxStages = 8;
constexpr int tma_consumer_warps(int tile_w) { return tile_w / kAnchorsPerWarp; }
constexpr int tma_threads(int tile_w) { return (tma_consumer_warps(tile_w) + 1) * 32; }  // + one producer warp

struct TmaGeom {
  HeadGeom g;
  int tiles_before[4];  // tiles of one stream before level l (prefix), [3] = tiles per stream
  int total_tiles;
  int stages;
  int tile_bytes;
  int evict_first;    // L2 evict-first hint on the tile loads
  int static_rounds;  // ring rounds with the static schedule (tile = blockIdx + k * grid) before tickets take over
  int trigger;        // release programmatic dependents (the next scan on the same stream) at once
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
  // bounded: a lost completion traps (launch error) instead of hanging the GPU
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int b,
                                              const uint64_t policy) {
  if (policy) {  // read-once data: L2 evict-first
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(b), "l"(policy)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(b)
        : "memory");
  }
}

// two horizontally adjacent anchors of one channel row: one 4-byte (16-bit heads) or 8-byte load,
// running maximum kept packed (HMNMX2 on bf16x2 / f16x2)
template <typename T>
struct Pair;
template <>
struct Pair<__nv_bfloat16> {
  using V = __nv_bfloat162;
  static __device__ __forceinline__ V lowest() { return __float2bfloat162_rn(-INFINITY); }
  static __device__ __forceinline__ V load(const __nv_bfloat16* p) { return *reinterpret_cast<const V*>(p); }
  static __device__ __forceinline__ V vmax(V a, V b) { return __hmax2(a, b); }
  static __device__ __forceinline__ float lo(V v) { return __low2float(v); }
  static __device__ __forceinline__ float hi(V v) { return __high2float(v); }
};
template <>
struct Pair<__half> {
  using V = __half2;
  static __device__ __forceinline__ V lowest() { return __float2half2_rn(-INFINITY); }
  static __device__ __forceinline__ V load(const __half* p) { return *reinterpret_cast<const V*>(p); }
  static __device__ __forceinline__ V vmax(V a, V b) { return __hmax2(a, b); }
  static __device__ __forceinline__ float lo(V v) { return __low2float(v); }
  static __device__ __forceinline__ float hi(V v) { return __high2float(v); }
};
template <>
struct Pair<float> {
  using V = float2;
  static __device__ __forceinline__ V lowest() { return make_float2(-INFINITY, -INFINITY); }
  static __device__ __forceinline__ V load(const float* p) { return *reinterpret_cast<const V*>(p); }
  static __device__ __forceinline__ V vmax(V a, V b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
  static __device__ __forceinline__ float lo(V v) { return v.x; }
  static __device__ __forceinline__ float hi(V v) { return v.y; }
};

// exact N1 of one anchor column for the lanes of its four class quarters: probability and index
// of the FIRST class attaining the maximum float32 sigmoid; lanes whose quarter cannot pass
// contribute (-1, INT_MAX).  `m` is the lane's maximum logit over its classes q, q+4, ...
// The quarter's class values are read in one unrolled sweep (independent shared-memory loads) that
// only notes which of them clear the gate - a handful at most; the sigmoid is evaluated for those,
// in ascending class order with a strict comparison, which is torch's max(1) on the sigmoid tensor.
template <typename T, bool NC80, int kTileW>
__device__ __forceinline__ void anchor_best(const T* cls_col, const int q, const int iters, const int nc, const float m,
                                            const float logit_gate, float* best, int* bc) {
  float sc = -1.f;
  int j = 0x7fffffff;
  if (m > logit_gate) {
    unsigned bits = 0u;
    if (NC80) {
#pragma unroll
      for (int i = 0; i < 20; ++i) bits |= (to_float(cls_col[(4 * i + q) * kTileW]) > logit_gate ? 1u : 0u) << i;
    } else {
      for (int i = 0; i < iters && i < 32; ++i)
        if (4 * i + q < nc) bits |= (to_float(cls_col[(4 * i + q) * kTileW]) > logit_gate ? 1u : 0u) << i;
    }
    while (bits) {
      const int i = __ffs(bits) - 1;
      bits &= bits - 1;
      const float p = sigmoidf_rn(to_float(cls_col[(4 * i + q) * kTileW]));
      if (p > sc) {
        sc = p;
        j = 4 * i + q;
      }
    }
    for (int i = 32; i < iters; ++i) {  // nc > 128: the classes the bitmask does not cover
      const int c = 4 * i + q;
      if (c < nc) {
        const float v = to_float(cls_col[c * kTileW]);
        if (v > logit_gate) {
          const float p = sigmoidf_rn(v);
          if (p > sc) {
            sc = p;
            j = c;
          }
        }
      }
    }
  }
#pragma unroll
  for (int d = 8; d <= 16; d <<= 1) {
    const float ob = __shfl_xor_sync(kFull, sc, d);
    const int oc = __shfl_xor_sync(kFull, j, d);
    if (ob > sc || (ob == sc && oc < j)) {
      sc = ob;
      j = oc;
    }
  }
  *best = sc;
  *bc = j;
}

// SPLIT: a tile holds the class planes only (the part every anchor needs); the 64 box channels are
// fetched per consumer warp - a 64 x 16-anchor sub-tile (one 32-byte sector per row for 16-bit heads),
// by a second TMA copy the warp issues itself - and only when its 16 anchors hold a candidate.  That
// copy is waited for one tile later (the warp scans the next tile meanwhile), so its latency stays
// off the ring.  Box bytes that no candidate needs never cross HBM.
struct TmaMaps {
  CUtensorMap tile[3];  // per level: the ring's tiles
  CUtensorMap box[3];   // per level: 64 box channels x 16 anchors (SPLIT only)
};

template <typename T, bool NC80, int kTileW, bool SPLIT>
__global__ void __launch_bounds__(tma_threads(kTileW)) decode_tma_kernel(const __grid_constant__ TmaMaps maps,
                                                                 const TmaGeom tg, const rtm_nms_params prm,
                                                                 const float logit_gate, const Workspace ws) {
  const CUtensorMap &map0 = maps.tile[0], &map1 = maps.tile[1], &map2 = maps.tile[2];
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ __align__(8) uint64_t box_bar[8][2];  // SPLIT: per consumer warp, two box sub-tiles in flight
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ int4 s_tile[kMaxStages];  // per stage: (stream, level, first anchor of the tile within the level, -) ; x < 0 = no more tiles
  __shared__ int s_next[kMaxStages];  // ticket drawn for the stage's next fill (producer lane only)
  using P = Pair<T>;
  constexpr int kConsumerWarps = kTileW / kAnchorsPerWarp;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = tg.stages;
  const int tps = tg.tiles_before[3], tb1 = tg.tiles_before[1], tb2 = tg.tiles_before[2];
  // Programmatic dependent launch: this grid may have been started while the PREVIOUS step's post
  // kernel (which releases its dependents as its first instruction) was still running.  Nothing here
  // reads or writes what that kernel touches: the head tensors are inputs, the candidate list goes
  // to another slot of the ring, the ticket counter of that slot was re-armed two steps ago.  The
  // step's own post kernel is an ordinary launch and starts after everything before it has finished.
  // The scan releases its own dependents at once as well: when the next launch on its stream is the
  // next step's scan (rtm_step_io.scan_async), that grid's CTAs move in as this one's retire.
  if (tg.trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    for (int w = 0; w < 8; ++w) {
      mbar_init(&box_bar[w][0], 1);
      mbar_init(&box_bar[w][1], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Tiles are handed out by a ticket counter (tile t = stream t / tps, tile t % tps of that stream),
  // so whichever CTAs are resident share the work evenly - also while the previous step's
  // post kernel still occupies part of the GPU.
  if (warp == kConsumerWarps) {
    // ===== producer warp: one elected lane keeps the ring full =====
    if (lane == 0) {
      uint64_t policy = 0;
      if (tg.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      auto issue = [&](int s, int t) {
        const int b = t / tps, r = t - b * tps;
        const int li = r >= tb2 ? 2 : (r >= tb1 ? 1 : 0);
        const int x = (r - (li == 2 ? tb2 : (li == 1 ? tb1 : 0))) * kTileW;
        s_tile[s] = make_int4(b, li, x, 0);
        mbar_expect_tx(&full_bar[s], tg.tile_bytes);
        tma_load_tile(tile_smem + static_cast<size_t>(s) * tg.tile_bytes, li == 0 ? &map0 : (li == 1 ? &map1 : &map2),
                      &full_bar[s], x, SPLIT ? kBoxCh : 0, b, policy);
      };
      // first round of the ring: tiles blockIdx + k * grid, no ticket needed; the tickets of the
      // second round are drawn meanwhile (all in flight together), later ones one ring cycle ahead
      const int nstatic = stages * tg.static_rounds;
      const int dyn0 = nstatic * static_cast<int>(gridDim.x);
      // (static_rounds = 0: the first round comes off the counter as well, `stages` tickets in one draw - a CTA
      // that only becomes resident late, e.g. behind another kernel's CTAs, then holds no tile of its own back)
      const int first = nstatic == 0 ? atomicAdd(ws.tile_counter, stages) : 0;
      bool done = false;
      for (int k = 0; k < stages && !done; ++k) {
        const int t = nstatic == 0 ? first + k : static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
        if (t < tg.total_tiles) {
          issue(k, t);
        } else {
          s_tile[k] = make_int4(-1, 0, 0, 0);
          mbar_arrive(&full_bar[k]);
          done = true;
        }
      }
      if (!done) {
        // tiles stages .. nstatic-1 of this CTA are static too; tickets are drawn for the ones after
        int tk[kMaxStages];
#pragma unroll
        for (int k = 0; k < kMaxStages; ++k)
          tk[k] = (k < stages && stages + k >= nstatic) ? atomicAdd(ws.tile_counter, 1) : 0;
#pragma unroll
        for (int k = 0; k < kMaxStages; ++k)
          if (k < stages)
            s_next[k] = stages + k < nstatic ? static_cast<int>(blockIdx.x) + (stages + k) * static_cast<int>(gridDim.x) : dyn0 + tk[k];
        int issued = 2 * stages;  // tiles of this CTA that have a source by now
        int s = 0, fill = 1, drawn = 0, drawn_for = -1;  // ticket in flight and the stage it is for
        while (true) {
          mbar_wait(&empty_bar[s], (fill - 1) & 1);
          if (drawn_for >= 0) s_next[drawn_for] = dyn0 + drawn;  // arrived while the ring drained
          const int t = s_next[s];
          if (t >= tg.total_tiles) {
            s_tile[s] = make_int4(-1, 0, 0, 0);
            mbar_arrive(&full_bar[s]);
            break;
          }
          issue(s, t);
          if (issued < nstatic) {
            s_next[s] = static_cast<int>(blockIdx.x) + issued * static_cast<int>(gridDim.x);
            drawn_for = -1;
          } else {
            drawn = atomicAdd(ws.tile_counter, 1);
            drawn_for = s;
          }
          ++issued;
          if (++s == stages) {
            s = 0;
            ++fill;
          }
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int pr = lane & 7, q = lane >> 3;          // anchor pair within the warp's 16, class quarter / DFL side
  const int col = warp * kAnchorsPerWarp + 2 * pr;  // first of the lane's two anchor columns in the tile
  const int nc = NC80 ? 80 : tg.g.num_classes;
  const int iters = (nc + 3) >> 2;  // quarter q scans classes q, q + 4, q + 8, ... (bank-conflict free)
  const int w0 = tg.g.lv[0].w, w1 = tg.g.lv[1].w, w2 = tg.g.lv[2].w;
  const int a1 = tg.g.lv[1].anchor0, a2 = tg.g.lv[2].anchor0;
  const int st0 = tg.g.lv[0].stride, st1 = tg.g.lv[1].stride, st2 = tg.g.lv[2].stride;
  const int hw0 = tg.g.lv[0].hw, hw1 = tg.g.lv[1].hw, hw2 = tg.g.lv[2].hw;
  uint8_t* mask_bytes = reinterpret_cast<uint8_t*>(ws.mask);

  // SPLIT: candidates of the previous tile whose box sub-tile is still in flight
  struct Pending {
    bool valid, cand0, cand1;
    float best0, best1;
    int bc0, bc1, b, li, pix;
  } pend;
  pend.valid = false;
  int box_cur = 0;
  unsigned box_phase = 0;  // bit i = parity to wait for on box_bar[warp][i]
  constexpr int kBoxElems = kBoxCh * kAnchorsPerWarp;  // elements of one box sub-tile
  T* box_buf = reinterpret_cast<T*>(tile_smem + static_cast<size_t>(stages) * tg.tile_bytes) + warp * 2 * kBoxElems;
  uint64_t box_policy = 0;
  if (SPLIT && tg.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(box_policy));
  auto level_of = [&](int li, int* lv_w, int* lv_stride, int* lv_anchor0) {
    *lv_w = li == 2 ? w2 : (li == 1 ? w1 : w0);
    *lv_stride = li == 2 ? st2 : (li == 1 ? st1 : st0);
    *lv_anchor0 = li == 2 ? a2 : (li == 1 ? a1 : 0);
  };
  // D1 for the lane's candidates from 16 bins x 4 sides (side q in this lane), then the store
  auto decode_and_store = [&](const float (&x0)[kRegMax], const float (&x1)[kRegMax], bool c0, bool c1, float bst0, float bst1,
                              int cl0, int cl1, int bb, int lli, int ppix) {
    const float d0 = c0 ? dfl_expectation(const_cast<float(&)[kRegMax]>(x0)) : 0.f;
    const float d1 = c1 ? dfl_expectation(const_cast<float(&)[kRegMax]>(x1)) : 0.f;
    const float t0 = __shfl_down_sync(kFull, d0, 8), r0 = __shfl_down_sync(kFull, d0, 16), b0 = __shfl_down_sync(kFull, d0, 24);
    const float t1 = __shfl_down_sync(kFull, d1, 8), r1 = __shfl_down_sync(kFull, d1, 16), b1 = __shfl_down_sync(kFull, d1, 24);
    if (q == 0) {
      int lv_w, lv_stride, lv_anchor0;
      level_of(lli, &lv_w, &lv_stride, &lv_anchor0);
      const int y = ppix / lv_w, x = ppix - y * lv_w;  // both anchors are in the same grid row (w is even)
      const float ay = static_cast<float>(y) + 0.5f, fs = static_cast<float>(lv_stride);
      if (c0)
        store_candidate(ws, bb, lv_anchor0 + ppix, dist_to_xyxy(d0, t0, r0, b0, static_cast<float>(x) + 0.5f, ay, fs, nullptr), bst0, cl0);
      if (c1)
        store_candidate(ws, bb, lv_anchor0 + ppix + 1,
                        dist_to_xyxy(d1, t1, r1, b1, static_cast<float>(x + 1) + 0.5f, ay, fs, nullptr), bst1, cl1);
    }
  };
  // SPLIT: the pending tile's box sub-tile has (or will soon have) arrived: finish its candidates
  auto finish_pending = [&](const int buf) {
    mbar_wait(&box_bar[warp][buf], (box_phase >> buf) & 1u);
    box_phase ^= 1u << buf;
    const T* bb = box_buf + buf * kBoxElems;
    float x0[kRegMax], x1[kRegMax];
    if (pend.cand0 || pend.cand1) {
#pragma unroll
      for (int k = 0; k < kRegMax; ++k) {
        const typename P::V v = P::load(bb + (q * kRegMax + k) * kAnchorsPerWarp + 2 * pr);
        x0[k] = P::lo(v);
        x1[k] = P::hi(v);
      }
    }
    __syncwarp();  // every lane has read the buffer before it can be refilled
    decode_and_store(x0, x1, pend.cand0, pend.cand1, pend.best0, pend.best1, pend.bc0, pend.bc1, pend.b, pend.li, pend.pix);
    pend.valid = false;
  };

  int s = 0, phase = 0;
  while (true) {
    mbar_wait(&full_bar[s], phase);
    int b, li, x0;
    asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, _}, [%3];" : "=r"(b), "=r"(li), "=r"(x0) : "r"(smem_u32(&s_tile[s])));
    if (b < 0) break;
    int lv_w, lv_stride, lv_anchor0;
    level_of(li, &lv_w, &lv_stride, &lv_anchor0);
    const int lv_hw = li == 2 ? hw2 : (li == 1 ? hw1 : hw0);
    const int pix = x0 + col;
    const T* tile = reinterpret_cast<const T*>(tile_smem + static_cast<size_t>(s) * tg.tile_bytes);
    const T* cls_rows = tile + (SPLIT ? 0 : kBoxCh) * kTileW;  // first class row of the tile
    const T* cls_col = cls_rows + q * kTileW + col;            // row of class q

    // ---- N1 gate: packed running maximum over this quarter's classes for both anchors ----
    typename P::V mv = P::lowest();
    if (NC80) {
#pragma unroll
      for (int i = 0; i < 20; ++i) mv = P::vmax(mv, P::load(cls_col + 4 * i * kTileW));
    } else {
      for (int i = 0; i < iters; ++i)
        if (4 * i + q < nc) mv = P::vmax(mv, P::load(cls_col + 4 * i * kTileW));
    }
    // anchors past the end of the level (zero-filled tail of the last tile) never pass; the test is
    // warp-uniform because every level holds a multiple of 16 anchors
    const bool in_level = pix < lv_hw;
    const float m0 = in_level ? P::lo(mv) : -INFINITY, m1 = in_level ? P::hi(mv) : -INFINITY;
    float am = fmaxf(m0, m1);
    am = fmaxf(am, __shfl_xor_sync(kFull, am, 8));
    am = fmaxf(am, __shfl_xor_sync(kFull, am, 16));

    bool cand0 = false, cand1 = false, released = false;
    if (__any_sync(kFull, am > logit_gate)) {
      // ---- exact N1 for the anchors that can pass, then D1 for the survivors ----
      float best0 = -1.f, best1 = -1.f;
      int bc0 = 0x7fffffff, bc1 = 0x7fffffff;
      if (NC80) {
        // both anchors of the lane in one sweep: a packed load yields the two class values of a row
        if (m0 > logit_gate || m1 > logit_gate) {
          unsigned bits0 = 0u, bits1 = 0u;
#pragma unroll
          for (int i = 0; i < 20; ++i) {
            const typename P::V v = P::load(cls_col + 4 * i * kTileW);
            bits0 |= (P::lo(v) > logit_gate ? 1u : 0u) << i;
            bits1 |= (P::hi(v) > logit_gate ? 1u : 0u) << i;
          }
          while (bits0) {  // ascending classes, strict >: the first maximum (torch's max(1) on the sigmoid tensor)
            const int i = __ffs(bits0) - 1;
            bits0 &= bits0 - 1;
            const float p = sigmoidf_rn(to_float(cls_col[4 * i * kTileW]));
            if (p > best0) {
              best0 = p;
              bc0 = 4 * i + q;
            }
          }
          while (bits1) {
            const int i = __ffs(bits1) - 1;
            bits1 &= bits1 - 1;
            const float p = sigmoidf_rn(to_float(cls_col[4 * i * kTileW + 1]));
            if (p > best1) {
              best1 = p;
              bc1 = 4 * i + q;
            }
          }
        }
#pragma unroll
        for (int d = 8; d <= 16; d <<= 1) {
          const float ob0 = __shfl_xor_sync(kFull, best0, d), ob1 = __shfl_xor_sync(kFull, best1, d);
          const int oc0 = __shfl_xor_sync(kFull, bc0, d), oc1 = __shfl_xor_sync(kFull, bc1, d);
          if (ob0 > best0 || (ob0 == best0 && oc0 < bc0)) {
            best0 = ob0;
            bc0 = oc0;
          }
          if (ob1 > best1 || (ob1 == best1 && oc1 < bc1)) {
            best1 = ob1;
            bc1 = oc1;
          }
        }
      } else {
        anchor_best<T, NC80, kTileW>(cls_rows + col, q, iters, nc, m0, logit_gate, &best0, &bc0);
        anchor_best<T, NC80, kTileW>(cls_rows + col + 1, q, iters, nc, m1, logit_gate, &best1, &bc1);
      }
      cand0 = best0 > prm.conf_thres && class_wanted(prm, bc0 & 255);
      cand1 = best1 > prm.conf_thres && class_wanted(prm, bc1 & 255);
      if (__any_sync(kFull, cand0 || cand1)) {
        if (SPLIT) {
          // the class planes are done with: hand the stage back, ask for this warp's box sub-tile, and
          // meanwhile finish the tile before this one
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&empty_bar[s]);
            mbar_expect_tx(&box_bar[warp][box_cur], kBoxElems * static_cast<int>(sizeof(T)));
            tma_load_tile(box_buf + box_cur * kBoxElems, li == 0 ? &maps.box[0] : (li == 1 ? &maps.box[1] : &maps.box[2]),
                          &box_bar[warp][box_cur], x0 + warp * kAnchorsPerWarp, 0, b, box_policy);
          }
          released = true;
          box_cur ^= 1;
                if (pend.valid) finish_pending(box_cur);  // requests alternate buffers: after the flip box_cur is the older one
          pend.valid = true;
          pend.cand0 = cand0;
          pend.cand1 = cand1;
          pend.best0 = best0;
          pend.best1 = best1;
          pend.bc0 = bc0;
          pend.bc1 = bc1;
          pend.b = b;
          pend.li = li;
          pend.pix = pix;
        } else {
          // side q of the lane's two anchors: both sets of 16 bins come out of shared memory first (a packed
          // load yields both anchors), then the stage is handed back to the producer and the arithmetic follows
          float x0v[kRegMax], x1v[kRegMax];
          if (cand0 || cand1) {
#pragma unroll
            for (int k = 0; k < kRegMax; ++k) {
              const typename P::V v = P::load(tile + (q * kRegMax + k) * kTileW + col);
              x0v[k] = P::lo(v);
              x1v[k] = P::hi(v);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);
          released = true;
          decode_and_store(x0v, x1v, cand0, cand1, best0, best1, bc0, bc1, b, li, pix);
        }
      }
    }
    // candidate bits of this warp's 16 anchors: two bytes of the stream's mask
    uint32_t even = __ballot_sync(kFull, cand0 && q == 0) & 0xffu, odd = __ballot_sync(kFull, cand1 && q == 0) & 0xffu;
    if (lane == 0 && in_level) {
      even = (even | (even << 4)) & 0x0f0fu;
      even = (even | (even << 2)) & 0x3333u;
      even = (even | (even << 1)) & 0x5555u;
      odd = (odd | (odd << 4)) & 0x0f0fu;
      odd = (odd | (odd << 2)) & 0x3333u;
      odd = (odd | (odd << 1)) & 0x5555u;
      *reinterpret_cast<uint16_t*>(mask_bytes + static_cast<size_t>(b) * ws.words * 4 + ((lv_anchor0 + pix) >> 3)) =
          static_cast<uint16_t>(even | (odd << 1));
    }
    if (!released) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);  // this warp is done with the stage
    }
    if (++s == stages) {
      s = 0;
      phase ^= 1;
    }
  }
  if (SPLIT && pend.valid) finish_pending(box_cur ^ 1);  // the most recent request
}

// ---------------------------------------------------------------------------------------
// decode_ldg (fallback)
// ---------------------------------------------------------------------------------------
constexpr int kVec = 8;  // anchors per thread = one byte of the candidate mask

// The warp decodes anchor `pix` of level `lv` of stream b: lane c handles channels c, c+32, ...
// Returns (uniformly) whether the anchor is a candidate; lane 0 stores it.
template <typename T>
__device__ __forceinline__ bool warp_decode_anchor(const T* __restrict__ base, const Level lv, int pix, int nc,
                                                   const rtm_nms_params& prm, const Workspace& ws, int b) {
  const int lane = threadIdx.x & 31;
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
    const float ob = __shfl_xor_sync(kFull, best, d);
    const int oc = __shfl_xor_sync(kFull, bc, d);
    if (ob > best || (ob == best && oc < bc)) {
      best = ob;
      bc = oc;
    }
  }
  if (!(best > prm.conf_thres)) return false;  // amax(1) > conf_thres, strict
  if (!class_wanted(prm, bc)) return false;    // `classes` filter acts on the arg-max class
  // box part: 4 lanes, one side each, sequential DFL (same operation order as the other kernels)
  float dist = 0.f;
  if (lane < 4) {
    float x[kRegMax];
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) x[k] = to_float(base[static_cast<size_t>(lane * kRegMax + k) * lv.hw + pix]);
    dist = dfl_expectation(x);
  }
  const float dt = __shfl_sync(kFull, dist, 1), dr = __shfl_sync(kFull, dist, 2), db = __shfl_sync(kFull, dist, 3);
  if (lane == 0) {
    const float ax = static_cast<float>(pix % lv.w) + 0.5f, ay = static_cast<float>(pix / lv.w) + 0.5f;
    store_candidate(ws, b, lv.anchor0 + pix, dist_to_xyxy(dist, dt, dr, db, ax, ay, static_cast<float>(lv.stride), nullptr),
                    best, bc);
  }
  return true;
}

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) decode_ldg_kernel(const HeadPtrs<T> heads, const HeadGeom g,
                                                             const rtm_nms_params prm, const float logit_gate,
                                                             const Workspace ws) {
  const int b = blockIdx.y;
  const int grp = blockIdx.x * THREADS + threadIdx.x;  // group of 8 consecutive anchors
  const int a0 = grp * kVec;
  int li = 0;
  if (a0 >= g.lv[1].anchor0) li = 1;
  if (a0 >= g.lv[2].anchor0) li = 2;
  const Level lv = g.lv[li];
  const bool in_range = a0 < g.num_anchors;
  const int pix0 = a0 - lv.anchor0;
  const int nc = g.num_classes;
  const T* base = heads.p[li] + static_cast<size_t>(b) * (kBoxCh + nc) * lv.hw;
  constexpr int kPer16 = 16 / sizeof(T);  // elements per 16-byte load

  float mx[kVec];
#pragma unroll
  for (int e = 0; e < kVec; ++e) mx[e] = -FLT_MAX;
  if (in_range) {
    const T* cls = base + static_cast<size_t>(kBoxCh) * lv.hw + pix0;
#pragma unroll 4
    for (int c = 0; c < nc; ++c) {
#pragma unroll
      for (int h = 0; h < kVec / kPer16; ++h) {
        const uint4 raw = *reinterpret_cast<const uint4*>(cls + static_cast<size_t>(c) * lv.hw + h * kPer16);
        const T* v = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int e = 0; e < kPer16; ++e) mx[h * kPer16 + e] = fmaxf(mx[h * kPer16 + e], to_float(v[e]));
      }
    }
  }
  unsigned pending = 0;
#pragma unroll
  for (int e = 0; e < kVec; ++e)
    if (in_range && mx[e] > logit_gate) pending |= 1u << e;

  unsigned mine = 0;  // candidate bits of this thread's 8 anchors
  unsigned lanes = __ballot_sync(kFull, pending != 0);
  while (lanes) {
    const int src = __ffs(lanes) - 1;
    const unsigned bits = __shfl_sync(kFull, pending, src);
    const int spix0 = __shfl_sync(kFull, pix0, src);
    const int sli = __shfl_sync(kFull, li, src);
    const Level slv = g.lv[sli];
    const T* sbase = heads.p[sli] + static_cast<size_t>(b) * (kBoxCh + nc) * slv.hw;
    unsigned rem = bits;
    while (rem) {
      const int e = __ffs(rem) - 1;
      rem &= rem - 1;
      const bool is_cand = warp_decode_anchor<T>(sbase, slv, spix0 + e, nc, prm, ws, b);
      if (is_cand && (threadIdx.x & 31) == src) mine |= 1u << e;
    }
    lanes &= lanes - 1;
  }
  if (in_range)
    reinterpret_cast<uint8_t*>(ws.mask + static_cast<size_t>(b) * ws.words)[a0 >> 3] = static_cast<uint8_t>(mine);
}

// ---------------------------------------------------------------------------------------
// decode_scan: streaming head scan without shared-memory staging.
//
// The class planes are the part of the head every anchor needs (80 of 144 channels); the 64 box
// channels matter only where a class passes the confidence test.  So the class planes are streamed
// once with 16-byte loads - four lanes share a group of 8 consecutive anchors, lane q taking
// classes q, q+4, ... - all of a lane's loads in flight together, the running maximum kept packed;
// and the box planes are read only for the 8-anchor groups that hold a candidate (their 16-byte
// pieces: sector-granular).  On the bench workload that is 2/3 of the head bytes.
// ---------------------------------------------------------------------------------------
constexpr int kScanThreads = 128;

template <typename T>
struct Vec16;  // eight anchors of one channel row: one (16-bit) or two (float) 16-byte loads
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int kLoads = 1;
  using M = __nv_bfloat162;
  static __device__ __forceinline__ M lowest() { return __float2bfloat162_rn(-INFINITY); }
  static __device__ __forceinline__ void fold(M (&m)[4], const uint4 (&v)[1]) {
    const M* p = reinterpret_cast<const M*>(&v[0]);
#pragma unroll
    for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], p[e]);
  }
  static __device__ __forceinline__ void unpack(const M (&m)[4], float (&f)[8]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      f[2 * e] = __low2float(m[e]);
      f[2 * e + 1] = __high2float(m[e]);
    }
  }
};
template <>
struct Vec16<__half> {
  static constexpr int kLoads = 1;
  using M = __half2;
  static __device__ __forceinline__ M lowest() { return __float2half2_rn(-INFINITY); }
  static __device__ __forceinline__ void fold(M (&m)[4], const uint4 (&v)[1]) {
    const M* p = reinterpret_cast<const M*>(&v[0]);
#pragma unroll
    for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], p[e]);
  }
  static __device__ __forceinline__ void unpack(const M (&m)[4], float (&f)[8]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      f[2 * e] = __low2float(m[e]);
      f[2 * e + 1] = __high2float(m[e]);
    }
  }
};
template <>
struct Vec16<float> {
  static constexpr int kLoads = 2;
  using M = float2;
  static __device__ __forceinline__ M lowest() { return make_float2(-INFINITY, -INFINITY); }
  static __device__ __forceinline__ void fold(M (&m)[4], const uint4 (&v)[2]) {
    const float* p = reinterpret_cast<const float*>(&v[0]);
#pragma unroll
    for (int e = 0; e < 4; ++e) m[e] = make_float2(fmaxf(m[e].x, p[2 * e]), fmaxf(m[e].y, p[2 * e + 1]));
  }
  static __device__ __forceinline__ void unpack(const M (&m)[4], float (&f)[8]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      f[2 * e] = m[e].x;
      f[2 * e + 1] = m[e].y;
    }
  }
};

__device__ __forceinline__ uint4 ld_stream16(const void* p) {  // read-once data: do not keep it in L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Shared memory of one warp of decode_scan_kernel: the rows of the current batch as loaded (so that the
// rare exact path can index them with run-time indices in compact loops - unrolled over 8 anchors x 20
// rows it overflowed the instruction cache: ncu showed `no_instructions` as the dominant stall and 174 us),
// and the running best (probability, class) of every (anchor, lane).
template <typename T, int BATCH>
struct ScanWarpSmem {
  uint4 rows[BATCH > kRegMax ? BATCH : kRegMax][32][Vec16<T>::kLoads];
  float best[8][32];
  int cls[8][32];
};

template <typename T, int BATCH>
__global__ void __launch_bounds__(kScanThreads) decode_scan_kernel(const HeadPtrs<T> heads, const HeadGeom g,
                                                                   const rtm_nms_params prm, const float logit_gate,
                                                                   const Workspace ws) {
  using V = Vec16<T>;
  constexpr int L = V::kLoads;
  extern __shared__ __align__(16) unsigned char scan_smem[];
  ScanWarpSmem<T, BATCH>& sm = reinterpret_cast<ScanWarpSmem<T, BATCH>*>(scan_smem)[threadIdx.x >> 5];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // see decode_tma_kernel
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, gl = lane & 7, q = lane >> 3;
  const int grp = (blockIdx.x * (kScanThreads / 32) + (threadIdx.x >> 5)) * 8 + gl;  // group of 8 consecutive anchors
  const int a0 = grp * kVec;
  const bool in_range = a0 < g.num_anchors;
  int li = 0;
  if (a0 >= g.lv[1].anchor0) li = 1;
  if (a0 >= g.lv[2].anchor0) li = 2;
  const int lv_hw = li == 2 ? g.lv[2].hw : (li == 1 ? g.lv[1].hw : g.lv[0].hw);
  const int lv_w = li == 2 ? g.lv[2].w : (li == 1 ? g.lv[1].w : g.lv[0].w);
  const int lv_stride = li == 2 ? g.lv[2].stride : (li == 1 ? g.lv[1].stride : g.lv[0].stride);
  const int lv_anchor0 = li == 2 ? g.lv[2].anchor0 : (li == 1 ? g.lv[1].anchor0 : 0);
  const int pix0 = a0 - lv_anchor0;
  const int nc = g.num_classes;
  const T* base = heads.p[li] + static_cast<size_t>(b) * (kBoxCh + nc) * lv_hw + pix0;
  const size_t hw = lv_hw;
  const T* my_rows = reinterpret_cast<const T*>(&sm.rows[0][lane][0]);  // element (row i, anchor j) at i * kRowElems + j
  constexpr int kRowElems = 32 * L * 16 / static_cast<int>(sizeof(T));

  // ---- N1: this lane's classes (q, q + 4, ...) over its 8 anchors.  A batch of rows is loaded (all
  //      loads in flight) and folded into a packed maximum; only if that maximum can pass the
  //      confidence test is the batch staged in shared memory and looked at element by element for
  //      the exact float32 sigmoid and the FIRST arg-max (ascending classes, strict >) ----
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm.best[j][lane] = -1.f;
    sm.cls[j][lane] = 0x7fffffff;
  }
  unsigned pass = 0;  // anchors of the group for which this lane's quarter cleared the gate at least once
  const int iters = (nc + 3) >> 2;
  for (int i0 = 0; i0 < iters; i0 += BATCH) {
    uint4 v[BATCH][L];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int c = 4 * (i0 + i) + q;
      const bool ok = in_range && c < nc;
#pragma unroll
      for (int l = 0; l < L; ++l)
        v[i][l] = ok ? ld_stream16(reinterpret_cast<const char*>(base + (kBoxCh + c) * hw) + 16 * l) : make_uint4(0, 0, 0, 0);
    }
    typename V::M mx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) mx[e] = V::lowest();
#pragma unroll
    for (int i = 0; i < BATCH; ++i)
      if (in_range && 4 * (i0 + i) + q < nc) V::fold(mx, v[i]);
    float am[8];
    V::unpack(mx, am);
    unsigned bpass = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (am[j] > logit_gate) bpass |= 1u << j;
    if (__any_sync(kFull, bpass != 0)) {  // rare per lane, common per warp: keep what follows compact
#pragma unroll
      for (int i = 0; i < BATCH; ++i)
#pragma unroll
        for (int l = 0; l < L; ++l) sm.rows[i][lane][l] = v[i][l];
      __syncwarp();
      pass |= bpass;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        if (!((bpass >> j) & 1u)) continue;
        float sc = sm.best[j][lane];
        int jc = sm.cls[j][lane];
#pragma unroll 1
        for (int i = 0; i < BATCH; ++i) {
          const int c = 4 * (i0 + i) + q;
          if (c >= nc) break;
          const float x = to_float(my_rows[i * kRowElems + j]);
          if (x > logit_gate) {
            const float p = sigmoidf_rn(x);
            if (p > sc) {
              sc = p;
              jc = c;
            }
          }
        }
        sm.best[j][lane] = sc;
        sm.cls[j][lane] = jc;
      }
      __syncwarp();
    }
  }

  unsigned cand = 0;  // candidate bits of the group's 8 anchors (identical in its four lanes)
  if (__any_sync(kFull, pass != 0)) {
    // ---- combine the four class quarters of every anchor; class filter on the arg-max class ----
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      if (!__any_sync(kFull, (pass >> j) & 1u)) continue;
      float sc = sm.best[j][lane];
      int jc = sm.cls[j][lane];
#pragma unroll
      for (int d = 8; d <= 16; d <<= 1) {
        const float ob = __shfl_xor_sync(kFull, sc, d);
        const int oc = __shfl_xor_sync(kFull, jc, d);
        if (ob > sc || (ob == sc && oc < jc)) {
          sc = ob;
          jc = oc;
        }
      }
      sm.best[j][lane] = sc;
      sm.cls[j][lane] = jc;
      if (sc > prm.conf_thres && class_wanted(prm, jc & 255)) cand |= 1u << j;
    }
    // ---- D1 for the candidates: lane q decodes side q of its group's candidate anchors ----
    if (__any_sync(kFull, cand != 0)) {
      if (cand) {
        uint4 bx[kRegMax][L];
#pragma unroll
        for (int k = 0; k < kRegMax; ++k)
#pragma unroll
          for (int l = 0; l < L; ++l)
            bx[k][l] = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(base + (q * kRegMax + k) * hw) + 16 * l);
#pragma unroll
        for (int k = 0; k < kRegMax; ++k)
#pragma unroll
          for (int l = 0; l < L; ++l) sm.rows[k][lane][l] = bx[k][l];
      }
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        if (!__any_sync(kFull, (cand >> j) & 1u)) continue;
        float dl = 0.f;
        if ((cand >> j) & 1u) {
          float x[kRegMax];
#pragma unroll
          for (int k = 0; k < kRegMax; ++k) x[k] = to_float(my_rows[k * kRowElems + j]);
          dl = dfl_expectation(x);
        }
        const float dt = __shfl_down_sync(kFull, dl, 8), dr = __shfl_down_sync(kFull, dl, 16), db = __shfl_down_sync(kFull, dl, 24);
        if (q == 0 && ((cand >> j) & 1u)) {
          const int pix = pix0 + j;
          const int y = pix / lv_w, x = pix - y * lv_w;
          store_candidate(ws, b, a0 + j,
                          dist_to_xyxy(dl, dt, dr, db, static_cast<float>(x) + 0.5f, static_cast<float>(y) + 0.5f,
                                       static_cast<float>(lv_stride), nullptr),
                          sm.best[j][lane], sm.cls[j][lane]);
        }
      }
    }
  }
  if (in_range && q == 0)
    reinterpret_cast<uint8_t*>(ws.mask + static_cast<size_t>(b) * ws.words)[a0 >> 3] = static_cast<uint8_t>(cand);
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
    float x[kRegMax];
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) x[k] = to_float(base[static_cast<size_t>(s * kRegMax + k) * lv.hw]);
    dist[s] = dfl_expectation(x);
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
__global__ void __launch_bounds__(256) pred_candidates_kernel(const float* __restrict__ pred, int A, int nc,
                                                              const rtm_nms_params prm, const Workspace ws) {
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;  // blockDim is a multiple of 32: a warp = one mask word
  bool cand = false;
  if (a < A) {
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
    cand = best > prm.conf_thres && class_wanted(prm, bc);
    if (cand) {
      const float cx = p[0], cy = p[A], w = p[2 * static_cast<size_t>(A)], h = p[3 * static_cast<size_t>(A)];
      const float dw = __fdiv_rn(w, 2.f), dh = __fdiv_rn(h, 2.f);
      store_candidate(ws, b, a, make_float4(__fsub_rn(cx, dw), __fsub_rn(cy, dh), __fadd_rn(cx, dw), __fadd_rn(cy, dh)),
                      best, bc);
    }
  }
  const uint32_t word = __ballot_sync(kFull, cand);
  if ((threadIdx.x & 31) == 0 && a < A) ws.mask[static_cast<size_t>(b) * ws.words + (a >> 5)] = word;
}

// ---------------------------------------------------------------------------------------
// nms
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const Workspace ws, const rtm_nms_params prm,
                                                          const float iou_gate, const NmsOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_keep[rtm::kMaxDetCap];
  __shared__ int s_scan[33];
  rtm::nms_stream<kNmsThreads>(ws, prm, iou_gate, out, blockIdx.x, smem_raw, s_keep, s_scan);
}

float logit_gate_for(float conf_thres) {
  // gate on the raw logit: anything whose sigmoid could exceed conf_thres passes (the exact
  // float32 test is repeated on the sigmoid itself)
  if (conf_thres <= 0.f) return -FLT_MAX;
  if (conf_thres >= 1.f) return FLT_MAX;
  const double lg = log(static_cast<double>(conf_thres) / (1.0 - static_cast<double>(conf_thres)));
  return static_cast<float>(lg - 1e-3 * (1.0 + fabs(lg)));
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
  RTM_REQUIRE(a0 < rtm::kMaxAnchors, "%d anchors exceed the supported %d", a0, rtm::kMaxAnchors - 1);
  for (int l = 0; l < 3; ++l)
    RTM_REQUIRE(g->lv[l].hw % kVec == 0, "level %d has %d anchors, not a multiple of %d", l, g->lv[l].hw, kVec);
  return RTM_OK;
}

int run_nms(const Workspace& ws, int B, const rtm_nms_params& prm, const NmsOut& out, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    RTM_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(rtm::kNmsSmemBytes)));
    configured = true;
  }
  {
    rtm::ProfileScope prof(RTM_K_NMS, stream);
    nms_kernel<<<B, kNmsThreads, rtm::kNmsSmemBytes, stream>>>(ws, prm, rtm::iou_gate_for(prm.iou_thres), out);
  }
  RTM_LAUNCH_CHECK("nms_kernel");
  return RTM_OK;
}

int check_common(const rtm_nms_params* p, int B, float* det_xyxy, float* det_conf, int32_t* det_cls,
                 int32_t* det_count, int det_stride, void* workspace) {
  RTM_REQUIRE(p, "null rtm_nms_params");
  RTM_REQUIRE(B > 0, "num_streams must be positive");
  RTM_REQUIRE(det_xyxy && det_conf && det_cls && det_count && workspace, "null output / workspace pointer");
  RTM_REQUIRE(p->max_det > 0 && p->max_det <= rtm::kMaxDetCap, "max_det %d out of range (1..%d)", p->max_det,
              rtm::kMaxDetCap);
  RTM_REQUIRE(det_stride >= p->max_det, "det_stride %d < max_det %d", det_stride, p->max_det);
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(det_xyxy) & 15) == 0, "det_xyxy must be 16-byte aligned");
  return RTM_OK;
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

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
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
template <typename T, int kTileW, bool SPLIT>
int launch_decode_tma_w(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B,
                        const rtm_nms_params& prm, const Workspace& ws, cudaStream_t stream) {
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) return 0;
  // a warp owns 16 consecutive anchors = two bytes of the stream's mask: every level must hold a
  // multiple of 16 anchors (so that groups never straddle levels) and have an even width (a lane's
  // two anchors share a grid row)
  for (int l = 0; l < 3; ++l)
    if (g.lv[l].hw % kAnchorsPerWarp != 0 || g.lv[l].w % 2 != 0) return 0;
  const int ch = kBoxCh + g.num_classes;
  const int tile_rows = SPLIT ? g.num_classes : ch;
  const void* ptrs[3] = {p3, p4, p5};
  // tensor maps are cached by what they describe (a caller cycles through a few sets of head buffers):
  // encoding six maps per step is host time the step does not have to spare
  struct MapKey {
    const void* p[3];
    int B, h, w, nc;
    bool operator==(const MapKey& o) const {
      return p[0] == o.p[0] && p[1] == o.p[1] && p[2] == o.p[2] && B == o.B && h == o.h && w == o.w && nc == o.nc;
    }
  };
  struct MapEntry {
    MapKey key;
    TmaMaps maps;
  };
  static std::vector<MapEntry> cache;  // per instantiation (element type, tile width, split)
  static size_t next_victim = 0;
  const MapKey key{{p3, p4, p5}, B, g.lv[0].h, g.lv[0].w, g.num_classes};
  const TmaMaps* cached = nullptr;
  for (const MapEntry& e : cache)
    if (e.key == key) {
      cached = &e.maps;
      break;
    }
  TmaMaps fresh;
  if (!cached) {
    TmaMaps& maps = fresh;
    static const int promo_env = env_int("RTM_TMA_L2PROMO", 3);  // 0 none, 1 64B, 2 128B, 3 256B
    const CUtensorMapL2promotion promo = promo_env == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : promo_env == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : promo_env == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                          : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    for (int l = 0; l < 3; ++l) {
      const cuuint64_t dims[3] = {static_cast<cuuint64_t>(g.lv[l].hw), static_cast<cuuint64_t>(ch), static_cast<cuuint64_t>(B)};
      const cuuint64_t strides[2] = {static_cast<cuuint64_t>(g.lv[l].hw) * sizeof(T),
                                     static_cast<cuuint64_t>(g.lv[l].hw) * ch * sizeof(T)};
      const cuuint32_t box[3] = {kTileW, static_cast<cuuint32_t>(tile_rows), 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      if (ch > 256 || (strides[0] & 15) != 0) return 0;
      CUresult r = encode(&maps.tile[l], tensor_map_dtype<T>(), 3, const_cast<void*>(ptrs[l]), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return 0;
      const cuuint32_t sub[3] = {kAnchorsPerWarp, kBoxCh, 1};  // a warp's box sub-tile (SPLIT); no L2 promotion: sector-sized rows
      r = encode(&maps.box[l], tensor_map_dtype<T>(), 3, const_cast<void*>(ptrs[l]), dims, strides, sub, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return 0;
    }
    if (cache.size() < 64) {
      cache.push_back(MapEntry{key, fresh});
    } else {
      cache[next_victim] = MapEntry{key, fresh};
      next_victim = (next_victim + 1) % cache.size();
    }
    cached = &fresh;
  }
  const TmaMaps& maps = *cached;
  TmaGeom tg;
  tg.g = g;
  tg.tiles_before[0] = 0;
  for (int l = 0; l < 3; ++l) tg.tiles_before[l + 1] = tg.tiles_before[l] + (g.lv[l].hw + kTileW - 1) / kTileW;
  tg.total_tiles = tg.tiles_before[3] * B;
  tg.tile_bytes = tile_rows * kTileW * static_cast<int>(sizeof(T));
  static const int static_env = env_int("RTM_TMA_STATIC_ROUNDS", 1);
  tg.static_rounds = static_env < 0 ? 0 : static_env;
  static const int evict_env = env_int("RTM_TMA_EVICT_FIRST", 1);
  tg.evict_first = evict_env;
  static const int trigger_env = env_int("RTM_SCAN_TRIGGER", 0);
  tg.trigger = trigger_env;
  if (tg.tile_bytes % 128 != 0) return 0;
  // ring depth and residency: as many tiles in flight per SM as fit (RTM_TMA_STAGES / RTM_TMA_CTAS override)
  static const int stages_env = env_int("RTM_TMA_STAGES", 0), ctas_env = env_int("RTM_TMA_CTAS", 0);
  const size_t smem_budget = 216 * 1024;
  const size_t box_smem = SPLIT ? static_cast<size_t>(kTileW / kAnchorsPerWarp) * 2 * kBoxCh * kAnchorsPerWarp * sizeof(T) : 0;
  int ctas_per_sm = ctas_env > 0 ? ctas_env : (SPLIT ? (sizeof(T) == 2 ? 3 : 2) : (tg.tile_bytes <= 24 * 1024 ? 3 : (tg.tile_bytes <= 40 * 1024 ? 2 : 1)));
  tg.stages = stages_env > 0 ? stages_env : (SPLIT ? (sizeof(T) == 2 ? 4 : 2) : (sizeof(T) == 2 ? 3 : 4));
  if (tg.stages > kMaxStages) tg.stages = kMaxStages;
  while (tg.stages > 1 && (static_cast<size_t>(tg.stages) * tg.tile_bytes + box_smem) * ctas_per_sm > smem_budget) --tg.stages;
  while (ctas_per_sm > 1 && (static_cast<size_t>(tg.stages) * tg.tile_bytes + box_smem) * ctas_per_sm > smem_budget) --ctas_per_sm;
  // RTM_TMA_SMEM_PAD_KB: unused shared memory on top of the ring (experiments: caps how many scan CTAs - of this and of
  // an overlapping scan - fit on an SM, i.e. how much room is left for another kernel's CTAs; 0 = none)
  static const int pad_env = env_int("RTM_TMA_SMEM_PAD_KB", 0);
  const size_t smem = static_cast<size_t>(tg.stages) * tg.tile_bytes + box_smem + static_cast<size_t>(pad_env > 0 ? pad_env : 0) * 1024;
  if (smem > 220 * 1024) return 0;
  static size_t configured = 0;
  if (smem > configured) {
    RTM_CUDA(cudaFuncSetAttribute(decode_tma_kernel<T, true, kTileW, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    RTM_CUDA(cudaFuncSetAttribute(decode_tma_kernel<T, false, kTileW, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  static const int grid_env = env_int("RTM_TMA_GRID", 0);  // experiments: any grid works with ticketed tiles
  const int grid = min(tg.total_tiles, grid_env > 0 ? grid_env : rtm::sm_count() * ctas_per_sm);
  {
    rtm::ProfileScope prof(RTM_K_DECODE, stream);
    // Programmatic dependent launch: when the kernel in front of this one on the stream is the
    // previous step's post kernel (which releases its dependents as soon as it starts), the scan
    // of this step runs beside it - it reads nothing that kernel writes (other workspace half).
    // Off while per-kernel profiling is on, so that each kernel is timed alone.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tma_threads(kTileW));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (rtm::pdl_enabled() && !rtm::g_profile_on) ? 1 : 0;
    const float gate = logit_gate_for(prm.conf_thres);
    if (g.num_classes == 80)
      RTM_CUDA(cudaLaunchKernelEx(&cfg, decode_tma_kernel<T, true, kTileW, SPLIT>, maps, tg, prm, gate, ws));
    else
      RTM_CUDA(cudaLaunchKernelEx(&cfg, decode_tma_kernel<T, false, kTileW, SPLIT>, maps, tg, prm, gate, ws));
  }
  RTM_LAUNCH_CHECK("decode_tma_kernel");
  return 1;
}

template <typename T>
int try_launch_decode_tma(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B,
                          const rtm_nms_params& prm, const Workspace& ws, cudaStream_t stream) {
  // anchors per tile (RTM_TMA_TILEW = 32 | 64 | 80 | 128 overrides): 80 when it divides every level
  // (640 x 640: 6400 / 1600 / 400; measured fastest, profiles/), else 64 / 32 with a masked tail
  static const int tile_env = env_int("RTM_TMA_TILEW", 0);
  bool div80 = true;
  for (int l = 0; l < 3; ++l) div80 = div80 && g.lv[l].hw % 80 == 0;
  const int tile_w = tile_env > 0 ? tile_env : (div80 ? 80 : (sizeof(T) == 2 ? 64 : 32));
  // RTM_TMA_SPLIT=1: class-plane tiles + per-warp box sub-tiles (80-wide tiles only).  Reads 27 % fewer bytes on
  // the bench workload but is slower there (37.4 vs 36.1 us alone: 40 % of the warp-tiles hold a candidate, and each
  // sub-tile is 64 requests of one sector); meant for sparse scenes, off by default
  static const int split_env = env_int("RTM_TMA_SPLIT", 0);
  switch (tile_w) {
    case 32:
      return launch_decode_tma_w<T, 32, false>(p3, p4, p5, g, B, prm, ws, stream);
    case 64:
      return launch_decode_tma_w<T, 64, false>(p3, p4, p5, g, B, prm, ws, stream);
    case 80:
      for (int l = 0; l < 3; ++l)
        if (g.lv[l].hw % 80 != 0) return 0;
      if (split_env) return launch_decode_tma_w<T, 80, true>(p3, p4, p5, g, B, prm, ws, stream);
      return launch_decode_tma_w<T, 80, false>(p3, p4, p5, g, B, prm, ws, stream);
    case 128:
      return launch_decode_tma_w<T, 128, false>(p3, p4, p5, g, B, prm, ws, stream);
    default:
      return 0;
  }
}

// which head scan to use: RTM_DECODE_IMPL = tma (default; falls back to scan where the tiling does
// not apply) | scan | ldg
int decode_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTM_DECODE_IMPL");
    v = !e ? 1 : (strcmp(e, "scan") == 0 ? 0 : (strcmp(e, "ldg") == 0 ? 2 : 1));
  }
  return v;
}

template <typename T>
int launch_decode(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B,
                  const rtm_nms_params& prm, const Workspace& ws, cudaStream_t stream) {
  HeadPtrs<T> heads{{static_cast<const T*>(p3), static_cast<const T*>(p4), static_cast<const T*>(p5)}};
  for (int l = 0; l < 3; ++l)
    RTM_REQUIRE((reinterpret_cast<uintptr_t>(heads.p[l]) & 15) == 0, "head level %d must be 16-byte aligned", l);
  const int impl = decode_impl();
  if (impl == 1) {
    const int tma = try_launch_decode_tma<T>(p3, p4, p5, g, B, prm, ws, stream);
    if (tma != 0) return tma < 0 ? tma : RTM_OK;
  }
  const float gate = logit_gate_for(prm.conf_thres);
  if (impl == 2) {
    constexpr int THREADS = 128;
    const int groups = g.num_anchors / kVec;
    dim3 grid((groups + THREADS - 1) / THREADS, B);
    {
      rtm::ProfileScope prof(RTM_K_DECODE, stream);
      decode_ldg_kernel<T, THREADS><<<grid, THREADS, 0, stream>>>(heads, g, prm, gate, ws);
    }
    RTM_LAUNCH_CHECK("decode_ldg_kernel");
    return RTM_OK;
  }
  // streaming scan: a warp takes 8 groups of 8 anchors
  const int groups = g.num_anchors / kVec;
  const int warps = (groups + 7) / 8;
  constexpr int BATCH = sizeof(T) == 2 ? 20 : 10;
  {
    rtm::ProfileScope prof(RTM_K_DECODE, stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((warps + kScanThreads / 32 - 1) / (kScanThreads / 32), B);
    cfg.blockDim = dim3(kScanThreads);
    cfg.dynamicSmemBytes = sizeof(ScanWarpSmem<T, BATCH>) * (kScanThreads / 32);
    static bool configured = false;
    if (!configured) {
      RTM_CUDA(cudaFuncSetAttribute(decode_scan_kernel<T, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(cfg.dynamicSmemBytes)));
      configured = true;
    }
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (rtm::pdl_enabled() && !rtm::g_profile_on) ? 1 : 0;
    RTM_CUDA(cudaLaunchKernelEx(&cfg, decode_scan_kernel<T, BATCH>, heads, g, prm, gate, ws));
  }
  RTM_LAUNCH_CHECK("decode_scan_kernel");
  return RTM_OK;
}

}  // namespace

extern "C" size_t rtm_nms_workspace_bytes(int32_t num_streams, int32_t num_anchors) {
  if (num_streams <= 0 || num_anchors <= 0) return 0;
  return workspace_layout(num_streams, num_anchors, nullptr, nullptr);
}

namespace {
std::unordered_map<const void*, int>& slot_table() {
  static std::unordered_map<const void*, int> t;
  return t;
}
}  // namespace

int rtm::next_scan_slot(const void* workspace) {
  const auto& t = slot_table();
  const auto it = t.find(workspace);
  return it == t.end() ? 0 : it->second;
}

int rtm::launch_decode_stage(const void* head_p3, const void* head_p4, const void* head_p5, int head_dtype,
                             int num_streams, int img_h, int img_w, const rtm_nms_params* params, void* workspace,
                             size_t workspace_bytes, Workspace* ws, cudaStream_t s) {
  RTM_REQUIRE(head_p3 && head_p4 && head_p5, "null head tensor");
  HeadGeom g;
  int rc = make_geom(img_h, img_w, params->num_classes, &g);
  if (rc) return rc;
  // Consecutive scans of one workspace take consecutive candidate-list slots (kept per workspace, so
  // that several stream batches, each with its own workspace and CUDA stream, can be interleaved).
  // The ticket counters must start at zero: the header is cleared the first time a workspace is seen;
  // afterwards every NMS stage leaves the counter of the slot it consumed at zero.
  std::unordered_map<const void*, int>& next_slot = slot_table();
  auto it = next_slot.find(workspace);
  if (it == next_slot.end()) {
    RTM_REQUIRE(workspace_bytes >= kWorkspaceHeader, "workspace too small");
    RTM_CUDA(cudaMemsetAsync(workspace, 0, kWorkspaceHeader, s));
    it = next_slot.emplace(workspace, 0).first;
  }
  const int half = it->second;
  it->second = (half + 1) % kCandSlots;
  const size_t need = workspace_layout(num_streams, g.num_anchors, static_cast<char*>(workspace), ws, half);
  RTM_REQUIRE(workspace_bytes >= need, "workspace has %zu bytes, %zu needed", workspace_bytes, need);
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  switch (head_dtype) {
    case RTM_F32:
      return launch_decode<float>(head_p3, head_p4, head_p5, g, num_streams, *params, *ws, s);
    case RTM_F16:
      return launch_decode<__half>(head_p3, head_p4, head_p5, g, num_streams, *params, *ws, s);
    case RTM_BF16:
      return launch_decode<__nv_bfloat16>(head_p3, head_p4, head_p5, g, num_streams, *params, *ws, s);
    default:
      RTM_REQUIRE(false, "unknown head_dtype %d", head_dtype);
  }
  return RTM_OK;
}

extern "C" int rtm_decode_nms(const void* head_p3, const void* head_p4, const void* head_p5,
                              int32_t head_dtype, int32_t num_streams, int32_t img_h, int32_t img_w,
                              const rtm_nms_params* params, const float* scale, float* det_xyxy,
                              float* det_conf, int32_t* det_cls, int32_t* det_anchor, int32_t* det_keep,
                              int32_t* det_count, int32_t det_stride, int32_t* status, void* workspace,
                              size_t workspace_bytes, rtm_cuda_stream stream) {
  int rc = check_common(params, num_streams, det_xyxy, det_conf, det_cls, det_count, det_stride, workspace);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace ws;
  rc = rtm::launch_decode_stage(head_p3, head_p4, head_p5, head_dtype, num_streams, img_h, img_w, params, workspace,
                                workspace_bytes, &ws, s);
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
  RTM_REQUIRE(num_anchors > 0 && num_anchors < rtm::kMaxAnchors, "rtm_nms_pred: num_anchors %d out of range", num_anchors);
  RTM_REQUIRE(params->num_classes > 0 && params->num_classes <= 256, "num_classes out of range");
  Workspace ws;
  const size_t need = workspace_layout(num_streams, num_anchors, static_cast<char*>(workspace), &ws);
  RTM_REQUIRE(workspace_bytes >= need, "rtm_nms_pred: workspace has %zu bytes, %zu needed", workspace_bytes, need);
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid((num_anchors + 255) / 256, num_streams);
  {
    rtm::ProfileScope prof(RTM_K_PRED, s);
    pred_candidates_kernel<<<grid, 256, 0, s>>>(pred, num_anchors, params->num_classes, *params, ws);
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
