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
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "decode_body.cuh"

namespace {

using rtm::kFull;
using rtm::NmsOut;
using rtm::Workspace;

using rtm::kRegMax;
using rtm::kBoxCh;
using rtm::Level;
using rtm::HeadGeom;
using rtm::TmaGeom;
using rtm::TmaMaps;
using rtm::to_float;
using rtm::sigmoidf_rn;
using rtm::class_wanted;
using rtm::dist_to_xyxy;
using rtm::dfl_expectation;
using rtm::store_candidate;
using rtm::kAnchorsPerWarp;
using rtm::kMaxStages;
using rtm::tma_threads;
constexpr int kNmsThreads = 512;

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Workspace = header + a ring of kCandSlots copies of the candidate interchange arrays + one set of
// spill arrays.  Consecutive head scans take consecutive slots, so the scan of a later step may already
// run while the post stage of an earlier step still reads its own slot.
// Header: per slot 128 bytes of counters - [0] tile tickets of the scan filling the slot (two-launch pipeline),
// and for the one-launch step (all running totals, never reset) [1] tiles of the slot's scans finished, [2] streams
// whose post stage has finished reading it, [3] post-stage tickets drawn - then the step kernel's per-workspace
// words: [0] the caller-stream mark, [64 .. 128) a ring of tile-ticket counters (one per launch, re-armed 32
// launches ahead), [128 + b] steps completed by stream b.
using rtm::kCandSlots;
constexpr size_t kSlotHeader = 128 * kCandSlots;
constexpr int kSyncWords = rtm::kSyncStreamSeq;  // words in front of the per-stream sequence numbers

size_t header_bytes(int B) { return align_up(kSlotHeader + sizeof(int) * (kSyncWords + static_cast<size_t>(B)), 256); }

size_t workspace_layout(int B, int A, char* base, Workspace* ws, int half = 0) {
  const int words = (A + 31) / 32, cap_p2 = next_pow2(A);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_head = take(header_bytes(B));
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
    ws->sync = reinterpret_cast<int*>(base + o_head + kSlotHeader);
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

template <typename T>
struct HeadPtrs {
  const T* p[3];
};

// ---------------------------------------------------------------------------------------
// decode_tma: the tiled head scan as a kernel of its own (decode_body.cuh holds the CTA's work)
// ---------------------------------------------------------------------------------------
template <typename T, bool NC80, int kTileW>
__global__ void __launch_bounds__(tma_threads(kTileW)) decode_tma_kernel(const __grid_constant__ TmaMaps maps,
                                                                 const TmaGeom tg, const rtm_nms_params prm,
                                                                 const float logit_gate, const Workspace ws) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ __align__(16) rtm::ScanCtl ctl;
  // Programmatic dependent launch (only attached when the launch before this one on the stream is the
  // library's own, see launch_decode_tma_w): nothing here reads or writes what the previous step's kernels
  // touch - the head tensors are inputs, the candidate list goes to another slot of the ring.
  if (tg.trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  rtm::tma_scan_cta<T, NC80, kTileW, 1>(maps, tg, prm, logit_gate, ws, rtm::ScanSync{nullptr, 0, nullptr},
                                         static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), tile_smem, &ctl);
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
  if (int rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(nms_kernel), rtm::kNmsSmemBytes)) return rc;
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

// Everything a launch of the tiled scan needs that can be derived from the head tensors: the (cached) tensor maps
// and the tile geometry.  Returns 1 when the tiling applies, 0 when the caller should fall back, < 0 on error.
// tg.stages is left to the caller.
template <typename T, int kTileW>
int plan_tma_scan(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B, TmaMaps* out_maps, TmaGeom* out_tg) {
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) return 0;
  // a warp owns 16 consecutive anchors = two bytes of the stream's mask: every level must hold a
  // multiple of 16 anchors (so that groups never straddle levels) and have an even width (a lane's
  // two anchors share a grid row)
  for (int l = 0; l < 3; ++l)
    if (g.lv[l].hw % kAnchorsPerWarp != 0 || g.lv[l].w % 2 != 0) return 0;
  const int ch = kBoxCh + g.num_classes;
  const void* ptrs[3] = {p3, p4, p5};
  // tensor maps are cached by what they describe (a caller cycles through a few sets of head buffers):
  // encoding three maps per step is host time the step does not have to spare
  struct MapKey {
    const void* p[3];
    int B, h, w, nc, dev;
    bool operator==(const MapKey& o) const {
      return p[0] == o.p[0] && p[1] == o.p[1] && p[2] == o.p[2] && B == o.B && h == o.h && w == o.w && nc == o.nc && dev == o.dev;
    }
  };
  struct MapEntry {
    MapKey key;
    TmaMaps maps;
  };
  static std::vector<MapEntry> cache;  // per instantiation (element type, tile width); guarded by the API mutex
  static size_t next_victim = 0;
  int dev = 0;
  RTM_CUDA(cudaGetDevice(&dev));
  const MapKey key{{p3, p4, p5}, B, g.lv[0].h, g.lv[0].w, g.num_classes, dev};
  const TmaMaps* cached = nullptr;
  for (const MapEntry& e : cache)
    if (e.key == key) {
      cached = &e.maps;
      break;
    }
  TmaMaps fresh;
  if (!cached) {
    static const int promo_env = env_int("RTM_TMA_L2PROMO", 3);  // 0 none, 1 64B, 2 128B, 3 256B
    const CUtensorMapL2promotion promo = promo_env == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : promo_env == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : promo_env == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                          : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    for (int l = 0; l < 3; ++l) {
      const cuuint64_t dims[3] = {static_cast<cuuint64_t>(g.lv[l].hw), static_cast<cuuint64_t>(ch), static_cast<cuuint64_t>(B)};
      const cuuint64_t strides[2] = {static_cast<cuuint64_t>(g.lv[l].hw) * sizeof(T),
                                     static_cast<cuuint64_t>(g.lv[l].hw) * ch * sizeof(T)};
      const cuuint32_t box[3] = {kTileW, static_cast<cuuint32_t>(ch), 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      if (ch > 256 || (strides[0] & 15) != 0) return 0;
      CUresult r = encode(&fresh.tile[l], tensor_map_dtype<T>(), 3, const_cast<void*>(ptrs[l]), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return 0;
      // the same tensor seen through a smaller box (lazy box rows): the class rows of a tile
      const cuuint32_t cls_box[3] = {kTileW, static_cast<cuuint32_t>(g.num_classes), 1};
      r = encode(&fresh.cls[l], tensor_map_dtype<T>(), 3, const_cast<void*>(ptrs[l]), dims, strides, cls_box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  *out_maps = *cached;
  TmaGeom& tg = *out_tg;
  tg.g = g;
  tg.tiles_before[0] = 0;
  for (int l = 0; l < 3; ++l) tg.tiles_before[l + 1] = tg.tiles_before[l] + (g.lv[l].hw + kTileW - 1) / kTileW;
  tg.total_tiles = tg.tiles_before[3] * B;
  tg.tile_bytes = ch * kTileW * static_cast<int>(sizeof(T));
  tg.cls_tile_bytes = g.num_classes * kTileW * static_cast<int>(sizeof(T));
  static const int static_env = env_int("RTM_TMA_STATIC_ROUNDS", 1);
  tg.static_rounds = static_env < 0 ? 0 : static_env;
  static const int evict_env = env_int("RTM_TMA_EVICT_FIRST", 1);
  tg.evict_first = evict_env;
  tg.trigger = 0;
  tg.l2_ahead = 0;
  tg.stages = 0;
  if (tg.tile_bytes % 128 != 0) return 0;
  return 1;
}

// returns 1 when the TMA path was launched, 0 when the caller should fall back, < 0 on error
template <typename T, int kTileW>
int launch_decode_tma_w(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B,
                        const rtm_nms_params& prm, const Workspace& ws, cudaStream_t stream, bool own_stream) {
  TmaMaps maps;
  TmaGeom tg;
  const int ok = plan_tma_scan<T, kTileW>(p3, p4, p5, g, B, &maps, &tg);
  if (ok <= 0) return ok;
  // RTM_SCAN_TRIGGER=1 (experiments): consecutive scans on the library's own stream release each other at once
  static const int trigger_env = env_int("RTM_SCAN_TRIGGER", 0);
  tg.trigger = trigger_env && own_stream;
  // ring depth and residency: as many tiles in flight per SM as fit (RTM_TMA_STAGES / RTM_TMA_CTAS override)
  static const int stages_env = env_int("RTM_TMA_STAGES", 0), ctas_env = env_int("RTM_TMA_CTAS", 0);
  const size_t smem_budget = 216 * 1024;
  int ctas_per_sm = ctas_env > 0 ? ctas_env : (tg.tile_bytes <= 24 * 1024 ? 3 : (tg.tile_bytes <= 40 * 1024 ? 2 : 1));
  tg.stages = stages_env > 0 ? stages_env : (sizeof(T) == 2 ? 3 : 4);
  if (tg.stages > kMaxStages) tg.stages = kMaxStages;
  while (tg.stages > 1 && static_cast<size_t>(tg.stages) * tg.tile_bytes * ctas_per_sm > smem_budget) --tg.stages;
  while (ctas_per_sm > 1 && static_cast<size_t>(tg.stages) * tg.tile_bytes * ctas_per_sm > smem_budget) --ctas_per_sm;
  // RTM_TMA_SMEM_PAD_KB: unused shared memory on top of the ring (experiments: caps how many scan CTAs fit on an SM)
  static const int pad_env = env_int("RTM_TMA_SMEM_PAD_KB", 0);
  const size_t smem = static_cast<size_t>(tg.stages) * tg.tile_bytes + static_cast<size_t>(pad_env > 0 ? pad_env : 0) * 1024;
  if (smem > 220 * 1024) return 0;
  if (int rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(decode_tma_kernel<T, true, kTileW>), smem)) return rc;
  if (int rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(decode_tma_kernel<T, false, kTileW>), smem)) return rc;
  static const int grid_env = env_int("RTM_TMA_GRID", 0);  // experiments: any grid works with ticketed tiles
  const int grid = min(tg.total_tiles, grid_env > 0 ? grid_env : rtm::sm_count() * ctas_per_sm);
  {
    rtm::ProfileScope prof(RTM_K_DECODE, stream);
    // Programmatic dependent launch ONLY on the library's own scan stream, where the launch in front of this one
    // is always the library's previous scan: a dependent that does not execute griddepcontrol.wait has no ordering
    // against its predecessor, so on a caller's stream (whose previous kernel may be the head producer, possibly
    // one that releases its dependents early) the scan is an ordinary launch.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tma_threads(kTileW));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (own_stream && rtm::pdl_enabled() && !rtm::g_profile_on) ? 1 : 0;
    const float gate = logit_gate_for(prm.conf_thres);
    if (g.num_classes == 80)
      RTM_CUDA(cudaLaunchKernelEx(&cfg, decode_tma_kernel<T, true, kTileW>, maps, tg, prm, gate, ws));
    else
      RTM_CUDA(cudaLaunchKernelEx(&cfg, decode_tma_kernel<T, false, kTileW>, maps, tg, prm, gate, ws));
  }
  RTM_LAUNCH_CHECK("decode_tma_kernel");
  return 1;
}

template <typename T>
int try_launch_decode_tma(const void* p3, const void* p4, const void* p5, const HeadGeom& g, int B,
                          const rtm_nms_params& prm, const Workspace& ws, cudaStream_t stream, bool own_stream) {
  // anchors per tile (RTM_TMA_TILEW = 32 | 64 | 80 | 128 overrides): 80 when it divides every level
  // (640 x 640: 6400 / 1600 / 400; measured fastest, profiles/), else 64 / 32 with a masked tail
  static const int tile_env = env_int("RTM_TMA_TILEW", 0);
  bool div80 = true;
  for (int l = 0; l < 3; ++l) div80 = div80 && g.lv[l].hw % 80 == 0;
  const int tile_w = tile_env > 0 ? tile_env : (div80 ? 80 : (sizeof(T) == 2 ? 64 : 32));
  switch (tile_w) {
    case 32:
      return launch_decode_tma_w<T, 32>(p3, p4, p5, g, B, prm, ws, stream, own_stream);
    case 64:
      return launch_decode_tma_w<T, 64>(p3, p4, p5, g, B, prm, ws, stream, own_stream);
    case 80:
      if (!div80) return 0;
      return launch_decode_tma_w<T, 80>(p3, p4, p5, g, B, prm, ws, stream, own_stream);
    case 128:
      return launch_decode_tma_w<T, 128>(p3, p4, p5, g, B, prm, ws, stream, own_stream);
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
                  const rtm_nms_params& prm, const Workspace& ws, cudaStream_t stream, bool own_stream) {
  HeadPtrs<T> heads{{static_cast<const T*>(p3), static_cast<const T*>(p4), static_cast<const T*>(p5)}};
  for (int l = 0; l < 3; ++l)
    RTM_REQUIRE((reinterpret_cast<uintptr_t>(heads.p[l]) & 15) == 0, "head level %d must be 16-byte aligned", l);
  const int impl = decode_impl();
  if (impl == 1) {
    const int tma = try_launch_decode_tma<T>(p3, p4, p5, g, B, prm, ws, stream, own_stream);
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
    if (int rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(decode_scan_kernel<T, BATCH>), cfg.dynamicSmemBytes)) return rc;
    cfg.stream = stream;  // an ordinary launch: ordered after whatever produced the head tensors on this stream
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
struct CtxKey {
  int device;
  const void* workspace;
  bool operator==(const CtxKey& o) const { return device == o.device && workspace == o.workspace; }
};
struct CtxKeyHash {
  size_t operator()(const CtxKey& k) const { return std::hash<const void*>()(k.workspace) ^ (static_cast<size_t>(k.device) << 1); }
};
std::unordered_map<CtxKey, rtm::WorkspaceCtx, CtxKeyHash>& ctx_table() {
  static std::unordered_map<CtxKey, rtm::WorkspaceCtx, CtxKeyHash> t;
  return t;
}
}  // namespace

std::recursive_mutex& rtm::api_mutex() {
  static std::recursive_mutex m;
  return m;
}

size_t rtm::workspace_header_bytes(int num_streams) { return header_bytes(num_streams); }

// Find or create the host bookkeeping of `workspace` on the current device.  The first time an address is seen
// its header (ticket / completion counters, per-stream sequence numbers) is cleared, synchronously - not on the
// per-frame path, and whichever stream uses the workspace first then finds it armed.  Callers hold api_mutex().
int rtm::workspace_ctx(void* workspace, size_t workspace_bytes, int num_streams, cudaStream_t s, rtm::WorkspaceCtx** out) {
  RTM_REQUIRE(workspace, "null workspace");
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  int dev = 0;
  RTM_CUDA(cudaGetDevice(&dev));
  auto& table = ctx_table();
  auto it = table.find(CtxKey{dev, workspace});
  if (it == table.end()) {
    const size_t head = header_bytes(num_streams);
    RTM_REQUIRE(workspace_bytes >= head, "workspace too small");
    RTM_CUDA(cudaMemsetAsync(workspace, 0, head, s));
    RTM_CUDA(cudaStreamSynchronize(s));
    it = table.emplace(CtxKey{dev, workspace}, rtm::WorkspaceCtx()).first;
    it->second.device = dev;
    it->second.header_streams = num_streams;
  } else if (num_streams > it->second.header_streams) {
    // a larger batch on the same address: the per-stream part of the header grows - start it from zero again
    rtm::WorkspaceCtx& c = it->second;
    if (c.stream) RTM_CUDA(cudaStreamSynchronize(c.stream));
    RTM_CUDA(cudaStreamSynchronize(s));
    RTM_CUDA(cudaMemsetAsync(workspace, 0, header_bytes(num_streams), s));
    RTM_CUDA(cudaStreamSynchronize(s));
    c.reset_counters();
    c.header_streams = num_streams;
  }
  *out = &it->second;
  return RTM_OK;
}

int rtm::workspace_streams(rtm::WorkspaceCtx* c) {
  if (c->stream) return RTM_OK;
  RTM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < rtm::kCandSlots; ++i) {
    RTM_CUDA(cudaEventCreateWithFlags(&c->scanned[i], cudaEventDisableTiming));
    RTM_CUDA(cudaEventCreateWithFlags(&c->consumed[i], cudaEventDisableTiming));
    RTM_CUDA(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
  }
  RTM_CUDA(cudaEventCreateWithFlags(&c->caller_mark, cudaEventDisableTiming));
  return RTM_OK;
}

// Takes the next slot of the workspace's candidate ring and describes it in *ws.
int rtm::take_scan_slot(rtm::WorkspaceCtx* c, void* workspace, size_t workspace_bytes, int num_streams, int num_anchors,
                        Workspace* ws) {
  const int slot = c->next_slot;
  c->next_slot = (slot + 1) % kCandSlots;
  const size_t need = workspace_layout(num_streams, num_anchors, static_cast<char*>(workspace), ws, slot);
  RTM_REQUIRE(workspace_bytes >= need, "workspace has %zu bytes, %zu needed", workspace_bytes, need);
  return RTM_OK;
}

extern "C" int rtm_workspace_release(void* workspace) {
  std::lock_guard<std::recursive_mutex> lock(rtm::api_mutex());
  int dev = 0;
  RTM_CUDA(cudaGetDevice(&dev));
  auto& table = ctx_table();
  auto it = table.find(CtxKey{dev, workspace});
  if (it == table.end()) return RTM_OK;
  rtm::WorkspaceCtx& c = it->second;
  if (c.stream) {
    RTM_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < rtm::kCandSlots; ++i) {
      cudaEventDestroy(c.scanned[i]);
      cudaEventDestroy(c.consumed[i]);
      cudaEventDestroy(c.done[i]);
    }
    cudaEventDestroy(c.caller_mark);
    cudaStreamDestroy(c.stream);
  }
  table.erase(it);
  return RTM_OK;
}

int rtm::geometry_for(int img_h, int img_w, int num_classes, HeadGeom* g) { return make_geom(img_h, img_w, num_classes, g); }

int rtm::launch_decode_stage(const void* head_p3, const void* head_p4, const void* head_p5, int head_dtype,
                             int num_streams, int img_h, int img_w, const rtm_nms_params* params, void* workspace,
                             size_t workspace_bytes, Workspace* ws, cudaStream_t s, rtm::WorkspaceCtx** out_ctx,
                             cudaStream_t scan_stream) {
  RTM_REQUIRE(head_p3 && head_p4 && head_p5, "null head tensor");
  HeadGeom g;
  int rc = make_geom(img_h, img_w, params->num_classes, &g);
  if (rc) return rc;
  // Consecutive scans of one workspace take consecutive candidate-list slots (kept per workspace, so
  // that several stream batches, each with its own workspace and CUDA stream, can be interleaved).
  // The ticket counters must start at zero: the header is cleared the first time a workspace is seen;
  // afterwards every NMS stage leaves the counter of the slot it consumed at zero.
  std::lock_guard<std::recursive_mutex> lock(rtm::api_mutex());
  rtm::WorkspaceCtx* ctx = nullptr;
  rc = rtm::workspace_ctx(workspace, workspace_bytes, num_streams, s, &ctx);
  if (rc) return rc;
  if (out_ctx) *out_ctx = ctx;
  rc = rtm::take_scan_slot(ctx, workspace, workspace_bytes, num_streams, g.num_anchors, ws);
  if (rc) return rc;
  const bool own = scan_stream != nullptr;
  cudaStream_t ls = own ? scan_stream : s;
  switch (head_dtype) {
    case RTM_F32:
      return launch_decode<float>(head_p3, head_p4, head_p5, g, num_streams, *params, *ws, ls, own);
    case RTM_F16:
      return launch_decode<__half>(head_p3, head_p4, head_p5, g, num_streams, *params, *ws, ls, own);
    case RTM_BF16:
      return launch_decode<__nv_bfloat16>(head_p3, head_p4, head_p5, g, num_streams, *params, *ws, ls, own);
    default:
      RTM_REQUIRE(false, "unknown head_dtype %d", head_dtype);
  }
  return RTM_OK;
}

int rtm::plan_tma_scan80(const void* p3, const void* p4, const void* p5, int head_dtype, int num_streams, int img_h,
                         int img_w, const rtm_nms_params* params, rtm::TmaScanPlan* plan) {
  RTM_REQUIRE(p3 && p4 && p5, "null head tensor");
  HeadGeom g;
  int rc = make_geom(img_h, img_w, params->num_classes, &g);
  if (rc) return rc;
  for (const void* p : {p3, p4, p5})
    RTM_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "head tensors must be 16-byte aligned");
  if (decode_impl() != 1 || env_int("RTM_TMA_TILEW", 80) != 80) return 0;
  if (rtm::kStepTileW == 80)
    for (int l = 0; l < 3; ++l)
      if (g.lv[l].hw % 80 != 0) return 0;
  plan->logit_gate = logit_gate_for(params->conf_thres);
  plan->nc80 = g.num_classes == 80;
  switch (head_dtype) {
    case RTM_F32:
      return plan_tma_scan<float, rtm::kStepTileW>(p3, p4, p5, g, num_streams, &plan->maps, &plan->tg);
    case RTM_F16:
      return plan_tma_scan<__half, rtm::kStepTileW>(p3, p4, p5, g, num_streams, &plan->maps, &plan->tg);
    case RTM_BF16:
      return plan_tma_scan<__nv_bfloat16, rtm::kStepTileW>(p3, p4, p5, g, num_streams, &plan->maps, &plan->tg);
    default:
      RTM_REQUIRE(false, "unknown head_dtype %d", head_dtype);
  }
  return 0;
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
