// D1 + N1 as a device function: the TMA head scan of one CTA (used by decode_tma_kernel in nms.cu and
// by the one-launch step kernel in post.cu).  See nms.cu for what is reproduced and DESIGN.md 5.1 for
// the numbers.
//
// A tile = all 64 + nc channels x kTileW consecutive anchors of one (stream, level), fetched by one TMA
// tensor copy into a ring of shared-memory stages (full / empty mbarriers).  A CTA holds GROUPS teams, each a
// ring of its own, a producer warp (one elected lane) that keeps it full and kTileW / 16 consumer warps that work on its tiles
// independently of each other, without a block barrier (one team per CTA in the stand-alone kernel, two in
// the step kernel, two of whose CTAs share an SM).  Tiles are handed out by a ticket counter after a static first
// ring round, so whichever CTAs are resident share the work evenly.
//
// LAZY (the step kernel's default): a tile is the nc CLASS rows only.  Consumer warps do N1 and push their candidates
// into a shared-memory queue (CandQueue); the block's other warps (cand_decoder_warp) read the 64 DFL values of every
// candidate from global memory and do D1 beside the scan - the box rows of anchors that are no candidates are never read.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>

#include "nms_body.cuh"

namespace rtm {

constexpr int kRegMax = 16;
constexpr int kBoxCh = 4 * kRegMax;  // 64

struct Level {
  int h, w, hw, stride;
  int anchor0;  // first anchor index of the level
};

struct HeadGeom {
  Level lv[3];
  int num_anchors;
  int num_classes;
};

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

// ---------------------------------------------------------------------------------------
// Tile width (anchors per tile) is a template parameter: 64 / 128 make every channel row of a
// 16-bit tile a whole number of 128-byte lines (one or two full-line L2 requests per row), 80
// divides 6400 / 1600 / 400 exactly but gives 160-byte rows that straddle lines.  Tiles that run
// past the end of a level are zero-filled by the TMA unit and their anchors masked out.
// ---------------------------------------------------------------------------------------
constexpr int kAnchorsPerWarp = 16;  // a lane owns 2 adjacent anchors x one class quarter
constexpr int kMaxStages = 12;
// anchors per tile of the one-launch step kernel (experiments: -DRTM_STEP_TILE_W=64 | 128; 80 divides every level of a
// 640 x 640 input exactly, other widths leave zero-filled tails that are masked out)
#ifndef RTM_STEP_TILE_W
#define RTM_STEP_TILE_W 80
#endif
constexpr int kStepTileW = RTM_STEP_TILE_W;
constexpr int tma_consumer_warps(int tile_w) { return tile_w / kAnchorsPerWarp; }
// threads of a scan CTA: GROUPS teams, each its consumer warps + a producer warp
constexpr int tma_threads(int tile_w, int groups = 1) { return groups * (tma_consumer_warps(tile_w) + 1) * 32; }

struct TmaGeom {
  HeadGeom g;
  int tiles_before[4];  // tiles of one stream before level l (prefix), [3] = tiles per stream
  int total_tiles;
  int stages;
  int tile_bytes;
  int evict_first;    // L2 evict-first hint on the tile loads
  int static_rounds;  // ring rounds with the static schedule (tile = blockIdx + k * grid) before tickets take over
  int trigger;        // release programmatic dependents (the next scan on the same stream) at once
  int l2_ahead;       // > 0: whoever draws ticket t also prefetches tile t + l2_ahead into L2 (a ring cycle of all teams)
  // lazy box rows (tma_scan_cta<..., LAZY = true>): the ring holds the class rows of a tile only; the 64 DFL values of
  // an anchor are read from global memory, by the block's decoder warps, for candidates only
  int cls_tile_bytes;  // class rows of a tile
};

struct TmaMaps {
  CUtensorMap tile[3];  // per level: the ring's tiles (all channels)
  CUtensorMap cls[3];   // per level: class rows only (box: tile width x num_classes x 1)
};

// LAZY: candidates on their way from the consumer warps (N1 done: stream, level, anchor, score, class) to the decoder
// warps (D1: the anchor's 64 DFL values from global memory, expectation, box, store).  A bounded ring in shared memory,
// many producers, many consumers, sequence numbers per slot: state == pos: free for the producer that drew position pos;
// pos + 1: filled; the decoder that drew pos reads it and sets pos + slots.
constexpr int kCandQueueSlots = 128;
struct CandQueue {
  int head, tail;
  int final_tail;  // < 0 while the scan runs; then the number of candidates pushed in all
  int state[kCandQueueSlots];
  int4 rec[kCandQueueSlots];  // stream | level << 24, anchor within the level, score bits, class
};

// What the one-launch step kernel adds around a scan (all null / zero for the stand-alone kernel):
// the scan of a candidate-ring slot may only start writing once the post stage that last read the slot
// is over, and it reports its own completion to the post CTAs of its launch.
struct ScanSync {
  const int* slot_free;  // counter the post stage bumps once per stream when it is done reading this slot
  int slot_free_target;  // value it must have reached before this scan may write the slot
  int* tiles_done;       // every scan CTA adds the number of tiles it has finished (once, at its end)
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
// the hardware may keep a waiting thread suspended for up to this long before try_wait returns false (it wakes at once
// when the phase completes): fewer polling instructions from waiting warps beside the post stage's warps
#ifndef RTM_MBAR_SUSPEND_NS
#define RTM_MBAR_SUSPEND_NS 2000
#endif
constexpr uint32_t kMbarSuspendNs = RTM_MBAR_SUSPEND_NS;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded: a lost completion traps (launch error) instead of hanging the GPU
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint32_t spin = 1; !mbar_try_wait(bar, parity); ++spin)
    if ((spin & 255u) == 0u) {  // 4 s, whatever a failed try costs
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// One thread waits until *p >= target (acquire).  Bounded in time: a dependency that never arrives traps
// (launch error) instead of hanging the GPU.
__device__ __forceinline__ void spin_until_ge(const int* p, const int target) {
  if (ld_acquire_gpu(p) - target >= 0) return;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_gpu(p) - target < 0) {
    __nanosleep(64);
    if (global_timer_ns() - t0 > 4000000000ull) __trap();
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
  // both halves = the largest value of the type that is <= x; 0xffff per half that is greater than the bound
  static __device__ __forceinline__ V floor_splat(float x) { return __bfloat162bfloat162(__float2bfloat16_rd(x)); }
  static __device__ __forceinline__ uint32_t gt_mask(V a, V b) { return __hgt2_mask(a, b); }
};
template <>
struct Pair<__half> {
  using V = __half2;
  static __device__ __forceinline__ V lowest() { return __float2half2_rn(-INFINITY); }
  static __device__ __forceinline__ V load(const __half* p) { return *reinterpret_cast<const V*>(p); }
  static __device__ __forceinline__ V vmax(V a, V b) { return __hmax2(a, b); }
  static __device__ __forceinline__ float lo(V v) { return __low2float(v); }
  static __device__ __forceinline__ float hi(V v) { return __high2float(v); }
  static __device__ __forceinline__ V floor_splat(float x) { return __half2half2(__float2half_rd(x)); }
  static __device__ __forceinline__ uint32_t gt_mask(V a, V b) { return __hgt2_mask(a, b); }
};
template <>
struct Pair<float> {
  using V = float2;
  static __device__ __forceinline__ V lowest() { return make_float2(-INFINITY, -INFINITY); }
  static __device__ __forceinline__ V load(const float* p) { return *reinterpret_cast<const V*>(p); }
  static __device__ __forceinline__ V vmax(V a, V b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
  static __device__ __forceinline__ float lo(V v) { return v.x; }
  static __device__ __forceinline__ float hi(V v) { return v.y; }
  static __device__ __forceinline__ V floor_splat(float x) { return make_float2(x, x); }
  static __device__ __forceinline__ uint32_t gt_mask(V a, V b) { return (a.x > b.x ? 0xffffu : 0u) | (a.y > b.y ? 0xffff0000u : 0u); }
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

// Shared-memory control block of a scan CTA (static; the tile ring itself is dynamic shared memory).
struct ScanCtl {
  uint64_t full_bar[kMaxStages];
  uint64_t empty_bar[kMaxStages];
  int4 tile[kMaxStages];  // per stage: (stream, level, first anchor of the tile within the level, anchors of the level from there on); x < 0 = no more tiles
  int mask_off[kMaxStages];  // per stage: byte offset of the tile's first anchor in the candidate mask
  int next[kMaxStages];   // ticket drawn for the stage's next fill (producer lane only)
  int issued[4];          // tiles each team's producer has handed to its consumers
  unsigned long long tl[4];  // RTM_TIMELINE builds: team 0's consumer wait ns, loop ns, tiles; its producer's wait ns
};

__device__ __forceinline__ int ld_volatile_shared(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_shared(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// all threads of the block, before a block barrier and before anyone pushes
__device__ __forceinline__ void cand_queue_init(CandQueue* cq) {
  for (int i = threadIdx.x; i < kCandQueueSlots; i += blockDim.x) cq->state[i] = i;
  if (threadIdx.x == 0) {
    cq->head = 0;
    cq->tail = 0;
    cq->final_tail = -1;
  }
}
__device__ __forceinline__ void cand_push(CandQueue* cq, int pos, int b, int li, int pix, float score, int cls) {
  const int slot = pos & (kCandQueueSlots - 1);
  for (uint32_t spin = 0; ld_volatile_shared(&cq->state[slot]) != pos; ++spin)  // a full ring: the decoders are behind
    if (spin > (1u << 26)) __trap();
  cq->rec[slot] = make_int4(b | (li << 24), pix, __float_as_int(score), cls);
  __threadfence_block();
  st_volatile_shared(&cq->state[slot], pos + 1);
}

// A decoder warp (LAZY): four lanes per candidate, one per box side; eight candidates per round.  Each lane reads its
// side's 16 DFL values of the anchor from global memory (one value per channel row: a 32-byte sector each; neighbouring
// candidates share sectors), forms the expectation with the function every decode path uses, and the side-0 lane
// assembles and stores the box.  Returns when the scan is over and the queue is empty.
template <typename T>
__device__ __forceinline__ void cand_decoder_warp(CandQueue* cq, const void* const (&head)[3], const TmaGeom& tg, const Workspace& ws) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, side = lane & 3;
  const int ch = kBoxCh + tg.g.num_classes;
  while (true) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&cq->head, 8);
    base = __shfl_sync(kFull, base, 0);
    const int pos = base + grp, slot = pos & (kCandQueueSlots - 1);
    bool have = false;
    int4 rec = make_int4(0, 0, 0, 0);
    if (side == 0) {
      for (uint32_t spin = 0;; ++spin) {
        if (ld_volatile_shared(&cq->state[slot]) == pos + 1) {
          have = true;
          break;
        }
        const int ft = ld_volatile_shared(&cq->final_tail);
        if (ft >= 0 && pos >= ft) break;
        __nanosleep(64);
        if (spin > (1u << 24)) __trap();
      }
      if (have) {
        __threadfence_block();
        rec = cq->rec[slot];
        st_volatile_shared(&cq->state[slot], pos + kCandQueueSlots);  // free for the producer one lap on
      }
    }
    have = __shfl_sync(kFull, have ? 1 : 0, grp * 4) != 0;
    rec.x = __shfl_sync(kFull, rec.x, grp * 4);
    rec.y = __shfl_sync(kFull, rec.y, grp * 4);
    rec.z = __shfl_sync(kFull, rec.z, grp * 4);
    rec.w = __shfl_sync(kFull, rec.w, grp * 4);
    if (!__any_sync(kFull, have)) break;  // every position of this round lies past the end: so will all later ones
    const int b = rec.x & 0xffffff, li = rec.x >> 24, pix = rec.y;
    float d = 0.f;
    if (have) {
      const int hw = tg.g.lv[li].hw;
      const T* src = static_cast<const T*>(head[li]) + (static_cast<size_t>(b) * ch + side * kRegMax) * hw + pix;
      float x[kRegMax];
#pragma unroll
      for (int k = 0; k < kRegMax; ++k) x[k] = to_float(__ldg(src + static_cast<size_t>(k) * hw));
      d = dfl_expectation(x);
    }
    const float t = __shfl_sync(kFull, d, grp * 4 + 1), r = __shfl_sync(kFull, d, grp * 4 + 2), bt = __shfl_sync(kFull, d, grp * 4 + 3);
    if (have && side == 0) {
      const int w = tg.g.lv[li].w;
      const int y = pix / w, xx = pix - y * w;
      store_candidate(ws, b, tg.g.lv[li].anchor0 + pix,
                      dist_to_xyxy(d, t, r, bt, static_cast<float>(xx) + 0.5f, static_cast<float>(y) + 0.5f,
                                   static_cast<float>(tg.g.lv[li].stride), nullptr),
                      __int_as_float(rec.z), rec.w);
    }
  }
}

// after the scan (and, LAZY, after the decoder warps have returned and a block barrier): one thread reports the CTA's tiles
__device__ __forceinline__ void scan_publish(const ScanSync& sync, const ScanCtl* ctl, int groups) {
  if (!sync.tiles_done) return;
  int done = 0;
  for (int g = 0; g < groups; ++g) done += ctl->issued[g];
  if (done) {
    __threadfence();
    atomicAdd(sync.tiles_done, done);
  }
}

// The scan of one CTA.  `cta` / `num_ctas`: this CTA's index among the scan CTAs of the launch and their
// number; the block has tma_threads(kTileW, GROUPS) threads, all of which must call it.  Ends with a block
// barrier after the last candidate store of the CTA; thread 0 then publishes how many tiles were done (sync.tiles_done).
template <typename T, bool NC80, int kTileW, int GROUPS, bool LAZY = false>
__device__ __forceinline__ void tma_scan_cta(const TmaMaps& maps, const TmaGeom& tg, const rtm_nms_params& prm,
                                             const float logit_gate, const Workspace& ws, const ScanSync& sync,
                                             const int cta, const int num_ctas, unsigned char* tile_smem, ScanCtl* ctl,
                                             CandQueue* cq = nullptr) {
  const CUtensorMap &map0 = LAZY ? maps.cls[0] : maps.tile[0], &map1 = LAZY ? maps.cls[1] : maps.tile[1],
                    &map2 = LAZY ? maps.cls[2] : maps.tile[2];
  using P = Pair<T>;
  constexpr int kTeamWarps = kTileW / kAnchorsPerWarp;
  constexpr int kConsumerWarps = GROUPS * kTeamWarps;
  const int stage_bytes = LAZY ? tg.cls_tile_bytes : tg.tile_bytes;  // what a ring stage holds
  uint64_t* full_bar = ctl->full_bar;
  uint64_t* empty_bar = ctl->empty_bar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = tg.stages;
  const int tps = tg.tiles_before[3], tb1 = tg.tiles_before[1], tb2 = tg.tiles_before[2];
  constexpr int kScanThreadsCta = GROUPS * (kTeamWarps + 1) * 32;  // the block may hold more threads: they do not come here
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kTeamWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kScanThreadsCta) : "memory");
  // (LAZY: the candidate queue was initialised by the caller, before a barrier of the whole block)

  // every team has a ring of its own: stages [team * spt, (team + 1) * spt), filled by its own producer warp
  // (GROUPS virtual CTAs side by side: the ticket round trip of a producer is hidden behind its team's tiles)
  const int spt = stages / GROUPS;
  if (warp >= kConsumerWarps) {
    // ===== producer warps: one elected lane each keeps the ring of its team full =====
    if (lane == 0) {
      const int team = warp - kConsumerWarps;
      const int s0 = team * spt;
      const int vcta = cta * GROUPS + team, vgrid = num_ctas * GROUPS;
      uint64_t policy = 0;
      if (tg.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map0)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map2)) : "memory");
      int handed = 0;  // tiles this producer has handed to its team
      const int mask_row_bytes = ws.words * 4;
      auto issue = [&](int s, int t) {
        ++handed;
        const int b = t / tps, r = t - b * tps;
        const int li = r >= tb2 ? 2 : (r >= tb1 ? 1 : 0);
        const int x = (r - (li == 2 ? tb2 : (li == 1 ? tb1 : 0))) * kTileW;
        ctl->mask_off[s0 + s] = b * mask_row_bytes + ((tg.g.lv[li].anchor0 + x) >> 3);
        ctl->tile[s0 + s] = make_int4(b, li, x, tg.g.lv[li].hw - x);
        mbar_expect_tx(&full_bar[s0 + s], stage_bytes);
        tma_load_tile(tile_smem + static_cast<size_t>(s0 + s) * stage_bytes, li == 0 ? &map0 : (li == 1 ? &map1 : &map2),
                      &full_bar[s0 + s], x, LAZY ? kBoxCh : 0, b, policy);
      };
#ifdef RTM_TIMELINE
      unsigned long long tl_empty = 0;
      auto wait_empty = [&](int s, int parity) {
        const unsigned long long t0 = global_timer_ns();
        mbar_wait(&empty_bar[s0 + s], parity);
        tl_empty += global_timer_ns() - t0;
      };
#else
      auto wait_empty = [&](int s, int parity) { mbar_wait(&empty_bar[s0 + s], parity); };
#endif
      // L2 prefetch of the tile that will be drawn `l2_ahead` tickets from now, by whichever team that will be: every
      // tile beyond the first ones is prefetched exactly once, about one ring cycle before its TMA load, which then
      // finds it in L2 - the ring's two stages per team cover an L2 hit, not a DRAM access under load
      auto prefetch_l2 = [&](int t) {
        if (tg.l2_ahead <= 0) return;
        t += tg.l2_ahead;
        if (t >= tg.total_tiles) return;
        const int b = t / tps, r = t - b * tps;
        const int li = r >= tb2 ? 2 : (r >= tb1 ? 1 : 0);
        const int x = (r - (li == 2 ? tb2 : (li == 1 ? tb1 : 0))) * kTileW;
        asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                         reinterpret_cast<uint64_t>(li == 0 ? &map0 : (li == 1 ? &map1 : &map2))),
                     "r"(x), "r"(LAZY ? kBoxCh : 0), "r"(b)
                     : "memory");
      };
      // first round of the ring: tiles vcta + k * vgrid, no ticket needed; the tickets of the
      // second round are drawn meanwhile (all in flight together), later ones one ring cycle ahead
      const int nstatic = spt * tg.static_rounds;
      const int dyn0 = nstatic * vgrid;
      // (static_rounds = 0: the first round comes off the counter as well, `spt` tickets in one draw - a CTA
      // that only becomes resident late, e.g. behind another kernel's CTAs, then holds no tile of its own back)
      const int first = nstatic == 0 ? atomicAdd(ws.tile_counter, spt) : 0;
      bool done = false;
      for (int k = 0; k < spt && !done; ++k) {
        const int t = nstatic == 0 ? first + k : vcta + k * vgrid;
        if (t < tg.total_tiles) {
          issue(k, t);
        } else {
          ctl->tile[s0 + k] = make_int4(-1, 0, 0, 0);
          mbar_arrive(&full_bar[s0 + k]);
          done = true;
        }
      }
      if (!done) {
        // tiles spt .. nstatic-1 of this team are static too; tickets are drawn for the ones after
        int tk[kMaxStages];
#pragma unroll
        for (int k = 0; k < kMaxStages; ++k)
          tk[k] = (k < spt && spt + k >= nstatic) ? atomicAdd(ws.tile_counter, 1) : 0;
#pragma unroll
        for (int k = 0; k < kMaxStages; ++k)
          if (k < spt) {
            ctl->next[s0 + k] = spt + k < nstatic ? vcta + (spt + k) * vgrid : dyn0 + tk[k];
            if (spt + k >= nstatic) prefetch_l2(dyn0 + tk[k]);
          }
        int issued = 2 * spt;  // tiles of this team that have a source by now
        int s = 0, fill = 1, drawn = 0, drawn_for = -1;  // ticket in flight and the stage it is for
        while (true) {
          wait_empty(s, (fill - 1) & 1);
          if (drawn_for >= 0) {
            ctl->next[s0 + drawn_for] = dyn0 + drawn;  // arrived while the ring drained
            prefetch_l2(dyn0 + drawn);
          }
          const int t = ctl->next[s0 + s];
          if (t >= tg.total_tiles) {
            ctl->tile[s0 + s] = make_int4(-1, 0, 0, 0);
            mbar_arrive(&full_bar[s0 + s]);
            break;
          }
          issue(s, t);
          if (issued < nstatic) {
            ctl->next[s0 + s] = vcta + issued * vgrid;
            drawn_for = -1;
          } else {
            drawn = atomicAdd(ws.tile_counter, 1);
            drawn_for = s;
          }
          ++issued;
          if (++s == spt) {
            s = 0;
            ++fill;
          }
        }
      }
      ctl->issued[team] = handed;
#ifdef RTM_TIMELINE
      if (team == 0) ctl->tl[3] = tl_empty;
#endif
    }
  } else {
    // ===== consumer warps =====
    const int team = warp / kTeamWarps, wit = warp - team * kTeamWarps;
    const int pr = lane & 7, q = lane >> 3;          // anchor pair within the warp's 16, class quarter / DFL side
    const int col = wit * kAnchorsPerWarp + 2 * pr;  // first of the lane's two anchor columns in the tile
    const int nc = NC80 ? 80 : tg.g.num_classes;
    const int iters = (nc + 3) >> 2;  // quarter q scans classes q, q + 4, q + 8, ... (bank-conflict free)
    const int w0 = tg.g.lv[0].w, w1 = tg.g.lv[1].w, w2 = tg.g.lv[2].w;
    const int a1 = tg.g.lv[1].anchor0, a2 = tg.g.lv[2].anchor0;
    const int st0 = tg.g.lv[0].stride, st1 = tg.g.lv[1].stride, st2 = tg.g.lv[2].stride;
    uint8_t* mask_bytes = reinterpret_cast<uint8_t*>(ws.mask);
    const typename P::V gate_pair = P::floor_splat(logit_gate);

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

    const int s0 = team * spt;  // the team's ring
    int s = s0, phase = 0;
    // the candidate slot this scan fills must have been let go by the post stage that last read it (a step that
    // is kCandSlots - 1 steps back: in practice never a wait, and the load travels while the first tiles do)
    if (sync.slot_free) {
      if (lane == 0) spin_until_ge(sync.slot_free, sync.slot_free_target);
      __syncwarp();
    }
#ifdef RTM_TIMELINE
    unsigned long long tl_wait = 0, tl_tiles = 0;
    const unsigned long long tl_begin = global_timer_ns();
#endif
    while (true) {
#ifdef RTM_TIMELINE
      const unsigned long long tl_t0 = global_timer_ns();
      mbar_wait(&full_bar[s], phase);
      tl_wait += global_timer_ns() - tl_t0;
      ++tl_tiles;
#else
      mbar_wait(&full_bar[s], phase);
#endif
      int b, li, x0, left;
      asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(b), "=r"(li), "=r"(x0), "=r"(left) : "r"(smem_u32(&ctl->tile[s])));
      if (b < 0) break;
      const int mask_off = ld_volatile_shared(&ctl->mask_off[s]) + 2 * wit;  // (before the stage can be handed back)
      const int pix = x0 + col;
      const T* tile = reinterpret_cast<const T*>(tile_smem + static_cast<size_t>(s) * stage_bytes);
      const T* cls_rows = LAZY ? tile : tile + kBoxCh * kTileW;  // first class row of the tile
      const T* cls_col = cls_rows + q * kTileW + col;  // row of class q

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
      const bool in_level = col < left;
      const float m0 = in_level ? P::lo(mv) : -INFINITY, m1 = in_level ? P::hi(mv) : -INFINITY;
      float am = fmaxf(m0, m1);
      am = fmaxf(am, __shfl_xor_sync(kFull, am, 8));
      am = fmaxf(am, __shfl_xor_sync(kFull, am, 16));

      bool cand0 = false, cand1 = false, released = false;
      uint32_t mask_bits = 0u;  // (lane 0)
      if (__any_sync(kFull, am > logit_gate)) {
        // ---- exact N1 for the anchors that can pass, then D1 for the survivors ----
        float best0 = -1.f, best1 = -1.f;
        int bc0 = 0x7fffffff, bc1 = 0x7fffffff;
        if (NC80) {
          // both anchors of the lane in one sweep: a packed load yields the two class values of a row
          if (m0 > logit_gate || m1 > logit_gate) {
            // which of the quarter's 20 classes can pass: a packed compare against the gate rounded DOWN to the element
            // type (a superset of the classes above the gate; what it adds cannot reach the confidence threshold)
            unsigned lo16 = 0u, hi4 = 0u;  // classes 0..15: anchor 0 in the low half, anchor 1 in the high half; 16..19 likewise
#pragma unroll
            for (int i = 0; i < 16; ++i)
              lo16 |= P::gt_mask(P::load(cls_col + 4 * i * kTileW), gate_pair) & ((1u << i) | (1u << (i + 16)));
#pragma unroll
            for (int i = 16; i < 20; ++i)
              hi4 |= P::gt_mask(P::load(cls_col + 4 * i * kTileW), gate_pair) & ((1u << (i - 16)) | (1u << i));
            unsigned bits0 = (lo16 & 0xffffu) | ((hi4 & 0xfu) << 16), bits1 = (lo16 >> 16) | (((hi4 >> 16) & 0xfu) << 16);
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
        const unsigned m0b = __ballot_sync(kFull, cand0 && q == 0) & 0xffu, m1b = __ballot_sync(kFull, cand1 && q == 0) & 0xffu;
        if (lane == 0) {  // candidate bits of the warp's 16 anchors, even and odd anchors interleaved
          uint32_t even = m0b, odd = m1b;
          even = (even | (even << 4)) & 0x0f0fu;
          even = (even | (even << 2)) & 0x3333u;
          even = (even | (even << 1)) & 0x5555u;
          odd = (odd | (odd << 4)) & 0x0f0fu;
          odd = (odd | (odd << 2)) & 0x3333u;
          odd = (odd | (odd << 1)) & 0x5555u;
          mask_bits = even | (odd << 1);
        }
        if (LAZY) {
          // hand the candidates to the decoder warps: the warp goes on with its next tile
          if (m0b | m1b) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&cq->tail, __popc(m0b) + __popc(m1b));
            base = __shfl_sync(kFull, base, 0);
            if (q == 0) {
              const unsigned below = (1u << lane) - 1u;
              if (cand0) cand_push(cq, base + __popc(m0b & below), b, li, pix, best0, bc0 & 255);
              if (cand1) cand_push(cq, base + __popc(m0b) + __popc(m1b & below), b, li, pix + 1, best1, bc1 & 255);
            }
          }
        } else if (__any_sync(kFull, cand0 || cand1)) {
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
      // candidate bits of this warp's 16 anchors: two bytes of the stream's mask (written whether or not any is set:
      // nothing clears the mask between frames)
      if (lane == 0 && in_level)
        *reinterpret_cast<uint16_t*>(mask_bytes + static_cast<uint32_t>(mask_off)) = static_cast<uint16_t>(mask_bits);
      if (!released) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);  // this warp is done with the stage
      }
      if (++s == s0 + spt) {
        s = s0;
        phase ^= 1;
      }
    }
#ifdef RTM_TIMELINE
    if (warp == 0 && lane == 0) {
      ctl->tl[0] = tl_wait;
      ctl->tl[1] = global_timer_ns() - tl_begin;
      ctl->tl[2] = tl_tiles - 1;
    }
#endif
  }
  // every candidate of this CTA's tiles is stored; publish that (the post stage of the launch waits for all tiles)
  asm volatile("bar.sync 1, %0;" ::"n"(kScanThreadsCta) : "memory");
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {  // the control block may be reused as ordinary shared memory
      mbar_inval(&full_bar[s]);
      mbar_inval(&empty_bar[s]);
    }
    if (LAZY) {  // every candidate of the CTA is in the queue (or decoded already): the decoder warps may run dry
      asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(&cq->final_tail)), "r"(*reinterpret_cast<volatile int*>(&cq->tail)) : "memory");
    }
    if (!LAZY && sync.tiles_done) {
      int done = 0;
      for (int g = 0; g < GROUPS; ++g) done += ctl->issued[g];
      if (done) {
        __threadfence();
        atomicAdd(sync.tiles_done, done);
      }
    }
  }
}

// Host side (nms.cu): everything a launch of the TMA scan needs, derived from the head tensors - cached tensor
// maps, tile geometry, ring depth.  Returns 1 when the tiled scan applies, 0 when it does not (shape, element
// type or driver without the tensor-map encoder), < 0 on error.  The ring depth (tg.stages, ring_bytes) is
// left to the caller; tiles are 80 anchors wide.
struct TmaScanPlan {
  TmaMaps maps;
  TmaGeom tg;
  float logit_gate;
  size_t ring_bytes;
  bool nc80;
};
int plan_tma_scan80(const void* p3, const void* p4, const void* p5, int head_dtype, int num_streams, int img_h, int img_w,
                    const rtm_nms_params* params, TmaScanPlan* plan);
int geometry_for(int img_h, int img_w, int num_classes, HeadGeom* g);

}  // namespace rtm
