// N2 + N3 for one stream, as a device function so that both the stand-alone NMS kernel and the
// fused post-backbone kernel can run it: candidate list -> torchvision-exact greedy NMS ->
// rescaled detections in score order.
//
// Candidate interchange format (written by the decode / filter kernels, read here):
//   mask  (B, ceil(A/32)) u32  bit a of stream b set = anchor a passed N1 (conf + class filter)
//   box   (B, A) float4        xyxy in letterbox pixels, valid where the bit is set
//   score (B, A) f32, cls (B, A) i32
// The list is dense by anchor, so it needs neither atomics nor a reset between frames, and its
// order IS the order of ultralytics' filtered tensor: the index torchvision.ops.nms returns for
// a survivor is simply its rank among the set bits.
#pragma once

#include <float.h>
#include <math.h>

#include "rtm_common.cuh"

namespace rtm {

constexpr float kMaxWh = 7680.f;    // ultralytics non_max_suppression max_wh
constexpr int kNmsSmemCand = 2048;  // candidates handled entirely in shared memory
constexpr int kIdxBits = 15;        // candidate rank (< 32768) in the low key bits
constexpr int kMaxAnchors = 1 << 15;
constexpr int kMaxDetCap = 1024;

struct Workspace {
  uint32_t* mask;  // (B, words)
  float4* box;     // (B, A)
  float* score;    // (B, A)
  int32_t* cls;    // (B, A)
  // spill arrays, used only when a stream has more than kNmsSmemCand candidates
  uint64_t* keys;   // (B, cap_p2)
  float4* ubox;     // (B, A)
  float4* sbox;     // (B, A)
  float* sarea;     // (B, A)
  int32_t* loc;     // (B, A)
  uint32_t* alive;  // (B, 2, words)
  int num_anchors, words, cap_p2;
};

struct NmsOut {
  const float* scale;  // (B, 5) gain, pad_x, pad_y, src_w, src_h or null
  float* xyxy;
  float* conf;
  int32_t* cls;
  int32_t* anchor;
  int32_t* keep;
  int32_t* count;
  int32_t stride;
  int32_t* status;
};

// smallest float32 g with (double)g > thr, so that `iou >= g` == `(double)iou > thr`
// (torchvision compares the float32 IoU with the double threshold)
inline float iou_gate_for(double thr) {
  const float f = static_cast<float>(thr);
  return static_cast<double>(f) > thr ? f : nextafterf(f, INFINITY);
}

// D1 + N1 only (defined in nms.cu): scans the head tensors and leaves the candidate list of every
// stream in `workspace`; *ws describes it for the NMS stage.
int launch_decode_stage(const void* head_p3, const void* head_p4, const void* head_p5, int head_dtype, int num_streams,
                        int img_h, int img_w, const rtm_nms_params* params, void* workspace, size_t workspace_bytes,
                        Workspace* ws, cudaStream_t stream);

// bytes of dynamic shared memory nms_stream needs
constexpr size_t kNmsSmemBytes = (sizeof(uint64_t) + 2 * sizeof(float4) + sizeof(float) + sizeof(int32_t)) * kNmsSmemCand +
                                 2 * sizeof(uint32_t) * (kNmsSmemCand / 32);

// IoU > thr decision of torchvision's CPU kernel: float32 IoU compared as a double with the
// double threshold == `iou >= gate` with gate = smallest float32 above thr.  The quotient is
// first formed with the fast reciprocal; only when it lands within 1e-5 relative of the gate
// is the IEEE division carried out, so the decision is always the exact one.
__device__ __forceinline__ bool suppresses(const float4 kb, const float ka, const float4 jb, const float ja,
                                           const float gate) {
  const float iw = fmaxf(0.f, __fsub_rn(fminf(kb.z, jb.z), fmaxf(kb.x, jb.x)));
  const float ih = fmaxf(0.f, __fsub_rn(fminf(kb.w, jb.w), fmaxf(kb.y, jb.y)));
  const float inter = __fmul_rn(iw, ih);
  const float den = __fsub_rn(__fadd_rn(ka, ja), inter);
  const float q = __fdividef(inter, den);
  if (fabsf(q - gate) > gate * 1e-5f) return q >= gate;  // NaN (0/0, zero-area pair) takes the exact path: false
  return __fdiv_rn(inter, den) >= gate;
}

// 64-bit bitonic sort of keys[0..P), P a power of two >= 32.  Strides >= 32 go through memory
// with one barrier each; strides < 32 are exchanged with warp shuffles.
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_keys(uint64_t* keys, const int P) {
  const int tid = threadIdx.x;
  auto register_phase = [&](int k_first, int k_last) {
    for (int i = tid; i < P; i += THREADS) {
      uint64_t x = keys[i];
      for (int k = k_first; k <= k_last; k <<= 1) {
        const bool up = (i & k) == 0;
        for (int j = min(k >> 1, 16); j >= 1; j >>= 1) {
          const uint64_t y = __shfl_xor_sync(kFull, x, j);
          const bool lower = (i & j) == 0;
          x = (lower == up) ? (x < y ? x : y) : (x < y ? y : x);
        }
      }
      keys[i] = x;
    }
    __syncthreads();
  };
  register_phase(2, min(P, 32));
  for (int k = 64; k <= P; k <<= 1) {
    for (int j = k >> 1; j >= 32; j >>= 1) {
      for (int i = tid; i < P; i += THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t x = keys[i], y = keys[ixj];
          if ((x > y) == ((i & k) == 0)) {
            keys[i] = y;
            keys[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
    register_phase(k, k);
  }
}

// Number of candidates of stream b (popcount of its mask).  All threads must call it.
template <int THREADS>
__device__ __forceinline__ int nms_count(const Workspace& ws, const int b, int* s_scan) {
  const int tid = threadIdx.x;
  const int A = ws.num_anchors, W = ws.words;
  const uint32_t* mask = ws.mask + static_cast<size_t>(b) * W;
  int n = 0;
  for (int w0 = 0; w0 < W; w0 += THREADS) {
    const int w = w0 + tid;
    uint32_t m = 0;
    if (w < W) {
      m = mask[w];
      const int valid = A - (w << 5);
      if (valid < 32) m &= (1u << valid) - 1u;
    }
    int tot;
    block_exclusive_sum(__popc(m), s_scan, &tot);
    n += tot;
  }
  return n;
}

// The NMS proper for a stream with n > 0 candidates.  IN_SMEM: working arrays live in `smem`
// (kNmsSmemBytes, n <= kNmsSmemCand); otherwise in the workspace spill arrays.
template <int THREADS, bool IN_SMEM, int kBatch>
__device__ __forceinline__ int nms_run(const Workspace& ws, const rtm_nms_params& prm, const float iou_gate,
                                       const NmsOut& out, const int b, const int n, unsigned char* smem,
                                       int* s_keep, int* s_scan) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = THREADS / 32;
  const int A = ws.num_anchors, W = ws.words;
  const size_t a0 = static_cast<size_t>(b) * A;
  const uint32_t* mask = ws.mask + static_cast<size_t>(b) * W;
  const int max_det = min(min(prm.max_det, out.stride), kMaxDetCap);
  uint64_t* keys;
  float4 *ubox, *sbox;
  float* sarea;
  int32_t* loc;
  uint32_t *alive0, *alive1;
  if (IN_SMEM) {
    keys = reinterpret_cast<uint64_t*>(smem);
    ubox = reinterpret_cast<float4*>(keys + kNmsSmemCand);
    sbox = ubox + kNmsSmemCand;
    sarea = reinterpret_cast<float*>(sbox + kNmsSmemCand);
    loc = reinterpret_cast<int32_t*>(sarea + kNmsSmemCand);
    alive0 = reinterpret_cast<uint32_t*>(loc + kNmsSmemCand);
    alive1 = alive0 + kNmsSmemCand / 32;
  } else {
    keys = ws.keys + static_cast<size_t>(b) * ws.cap_p2;
    ubox = ws.ubox + a0;
    sbox = ws.sbox + a0;
    sarea = ws.sarea + a0;
    loc = ws.loc + a0;
    alive0 = ws.alive + static_cast<size_t>(b) * 2 * W;
    alive1 = alive0 + W;
  }

  // ---- expand the set bits into loc[rank] = anchor ----
  int base = 0;
  for (int w0 = 0; w0 < W; w0 += THREADS) {
    const int w = w0 + tid;
    uint32_t m = 0;
    if (w < W) {
      m = mask[w];
      const int valid = A - (w << 5);
      if (valid < 32) m &= (1u << valid) - 1u;
    }
    int tot;
    int r = base + block_exclusive_sum(__popc(m), s_scan, &tot);
    while (m) {
      loc[r++] = (w << 5) + __ffs(m) - 1;
      m &= m - 1;
    }
    base += tot;
  }
  __syncthreads();

  // ---- stage candidates (one global round trip) and build the sort keys: descending score,
  //      then ascending rank (= torchvision's stable sort of the filtered list) ----
  int P = 32;
  while (P < n) P <<= 1;
  for (int i = tid; i < P; i += THREADS) {
    uint64_t key = ~0ull;
    if (i < n) {
      const int an = loc[i];
      const float sc = ws.score[a0 + an];
      ubox[i] = ws.box[a0 + an];
      loc[i] = an | (ws.cls[a0 + an] << 16);
      key = (static_cast<uint64_t>(~float_orderable(sc)) << 32) | static_cast<uint32_t>(i);
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_keys<THREADS>(keys, P);

  // ---- boxes in sorted order with the class offset added in float32, areas as torchvision ----
  const int words = (n + 31) >> 5;
  for (int i = tid; i < n; i += THREADS) {
    const int r = static_cast<int>(keys[i] & ((1u << kIdxBits) - 1u));
    float4 bx = ubox[r];
    const float off = prm.agnostic ? 0.f : __fmul_rn(static_cast<float>(loc[r] >> 16), kMaxWh);
    bx.x = __fadd_rn(bx.x, off);
    bx.y = __fadd_rn(bx.y, off);
    bx.z = __fadd_rn(bx.z, off);
    bx.w = __fadd_rn(bx.w, off);
    sbox[i] = bx;
    sarea[i] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
  }
  for (int w = tid; w < words; w += THREADS) {
    const int rem = n - (w << 5);
    alive0[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
  }
  __syncthreads();

  int kept = 0;
  if (IN_SMEM) {
    // ---- greedy scan, register-resident: lane (warp, lane) owns candidates j = tid + k*THREADS
    //      (word k*kWarps + warp of the alive bitmask); each barrier round settles the first
    //      kBatch alive candidates exactly as the serial scan would: the first is a survivor, each
    //      next one survives unless an earlier survivor of the batch suppresses it, and every
    //      other candidate is tested against the batch's survivors ----
    constexpr int KMAX = kNmsSmemCand / THREADS;
    float4 mybox[KMAX];
    float myarea[KMAX];
    bool myalive[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int j = tid + k * THREADS;
      myalive[k] = j < n;
      mybox[k] = myalive[k] ? sbox[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      myarea[k] = myalive[k] ? sarea[j] : 0.f;
    }
    uint32_t* cur = alive0;
    uint32_t* nxt = alive1;
    int w0 = 0;
    while (kept < max_det) {
      // first kBatch alive candidates in score order (uniform across the block)
      int c[kBatch], nb = 0;
      {
        int w = w0;
        uint32_t word = w < words ? cur[w] : 0u;
        while (w < words && nb < kBatch) {
          if (word == 0u) {
            ++w;
            if (nb == 0) w0 = w;
            word = w < words ? cur[w] : 0u;
          } else {
            c[nb++] = (w << 5) + __ffs(word) - 1;
            word &= word - 1;
          }
        }
      }
      if (nb == 0) break;
      float4 bb[kBatch];
      float ba[kBatch];
      bool surv[kBatch];
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        const int ci = i < nb ? c[i] : c[0];
        bb[i] = sbox[ci];
        ba[i] = sarea[ci];
        surv[i] = i < nb;
      }
#pragma unroll
      for (int i = 1; i < kBatch; ++i)
#pragma unroll
        for (int p2 = 0; p2 < i; ++p2)
          if (surv[i] && surv[p2] && suppresses(bb[p2], ba[p2], bb[i], ba[i], iou_gate)) surv[i] = false;
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        if (surv[i]) {
          if (kept < max_det) {
            if (tid == 0) s_keep[kept] = c[i];
            ++kept;
          } else {
            surv[i] = false;  // beyond max_det: never reported (the scan stops after this round)
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int j = tid + k * THREADS;
        bool alive = myalive[k];
        if (alive) {
#pragma unroll
          for (int i = 0; i < kBatch; ++i) {
            if (i < nb && (j == c[i] || (surv[i] && suppresses(bb[i], ba[i], mybox[k], myarea[k], iou_gate)))) alive = false;
          }
        }
        myalive[k] = alive;
        const uint32_t nw = __ballot_sync(kFull, alive);
        if (lane == 0 && k * kWarps + warp < words) nxt[k * kWarps + warp] = nw;
      }
      __syncthreads();
      uint32_t* t = cur;
      cur = nxt;
      nxt = t;
    }
  } else {
  // ---- greedy scan, one survivor per iteration; alive bits rebuilt by ballot (ping-pong) ----
  uint32_t* cur = alive0;
  uint32_t* nxt = alive1;
  int w0 = 0;
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
      if (alive) alive = !suppresses(kb, ka, sbox[j], sarea[j], iou_gate);
      const uint32_t nw = __ballot_sync(kFull, alive);
      if (lane == 0) nxt[w] = nw;
    }
    __syncthreads();
    uint32_t* t = cur;
    cur = nxt;
    nxt = t;
  }
  }
  __syncthreads();

  // ---- survivors in score order: original box -> scale_boxes -> clip ----
  float gain = 1.f, padx = 0.f, pady = 0.f, sw = 0.f, sh = 0.f;
  if (out.scale) {
    gain = out.scale[b * 5 + 0];
    padx = out.scale[b * 5 + 1];
    pady = out.scale[b * 5 + 2];
    sw = out.scale[b * 5 + 3];
    sh = out.scale[b * 5 + 4];
  }
  const size_t o0 = static_cast<size_t>(b) * out.stride;
  for (int o = tid; o < kept; o += THREADS) {
    const uint64_t key = keys[s_keep[o]];
    const int r = static_cast<int>(key & ((1u << kIdxBits) - 1u));
    float4 bx = ubox[r];
    if (out.scale) {
      bx.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.x, padx), gain), 0.f), sw);
      bx.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.y, pady), gain), 0.f), sh);
      bx.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.z, padx), gain), 0.f), sw);
      bx.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.w, pady), gain), 0.f), sh);
    }
    reinterpret_cast<float4*>(out.xyxy)[o0 + o] = bx;
    out.conf[o0 + o] = float_from_orderable(~static_cast<uint32_t>(key >> 32));
    const int meta = loc[r];
    out.cls[o0 + o] = meta >> 16;
    if (out.anchor) out.anchor[o0 + o] = meta & 0xffff;
    if (out.keep) out.keep[o0 + o] = r;  // torchvision's index: rank in the filtered list
  }
  if (tid == 0) out.count[b] = kept;
  __syncthreads();
  return kept;
}

// One stream.  `smem` = kNmsSmemBytes of dynamic shared memory, `s_keep` = kMaxDetCap ints,
// `s_scan` = 33 ints.  All THREADS threads of the block must call it.  On return (after a
// trailing __syncthreads) s_keep[0..kept) and the output slabs are written; returns kept.
template <int THREADS, int kBatch = 1>
__device__ __forceinline__ int nms_stream(const Workspace& ws, const rtm_nms_params& prm, const float iou_gate,
                                          const NmsOut& out, const int b, unsigned char* smem, int* s_keep,
                                          int* s_scan) {
  const int n = nms_count<THREADS>(ws, b, s_scan);
  if (n == 0) {
    if (threadIdx.x == 0) out.count[b] = 0;
    __syncthreads();
    return 0;
  }
  if (n <= kNmsSmemCand) return nms_run<THREADS, true, kBatch>(ws, prm, iou_gate, out, b, n, smem, s_keep, s_scan);
  return nms_run<THREADS, false, 1>(ws, prm, iou_gate, out, b, n, smem, s_keep, s_scan);
}

}  // namespace rtm
