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

#include <mutex>

#include "rtm_common.cuh"

namespace rtm {

constexpr float kMaxWh = 7680.f;    // ultralytics non_max_suppression max_wh
#ifndef RTM_NMS_SMEM_CAND
#define RTM_NMS_SMEM_CAND 1024  // streams with more candidates take the spill arrays of the workspace
#endif
constexpr int kNmsSmemCand = RTM_NMS_SMEM_CAND;  // candidates handled entirely in shared memory
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
  // tile ticket counter of the head scan that filled this candidate list; the NMS stage that
  // consumes the list zeroes it again (the workspace starts zero-filled)
  int* tile_counter;  // [1]: scan CTAs done, [2]: streams done reading the slot (one-launch step, post.cu)
  int* sync;          // step kernel: [0] caller-stream mark, [64 + launch % 64] tile tickets, [128 + b] steps completed by stream b
  int num_anchors, words, cap_p2;
  int slot;  // which slot of the candidate ring this is (host-side bookkeeping)
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

// Ring of candidate-list slots in the workspace: consecutive head scans fill consecutive slots.  Three would do
// for the overlap of a scan with the previous step's post kernel; the ring is deeper so that in scan_async mode
// the "slot free again" ordering can be enqueued for kSlotWaitEvery scans at a time instead of before every scan
// (an event wait between two scans keeps the second from being launched ahead, ~1.6 us per step when measured).
constexpr int kCandSlots = 8;
constexpr int kSlotWaitEvery = 4;  // must divide kCandSlots; kCandSlots - kSlotWaitEvery steps of run-ahead remain
constexpr int kSyncTicketRing = 64;   // ws.sync[kSyncTicketRing + (launch % 64)]: tile tickets of a step kernel launch
constexpr int kSyncStreamSeq = 128;   // ws.sync[kSyncStreamSeq + b]: steps completed by stream b

// Host bookkeeping of one workspace (nms.cu), keyed by (device, address); every entry point that touches it
// holds api_mutex().  rtm_workspace_release drops the entry together with the stream and events it owns.
struct WorkspaceCtx {
  int device = 0;
  int header_streams = 0;  // streams the cleared header was sized for
  int next_slot = 0;       // slot of the candidate ring the next scan takes
  cudaStream_t stream = nullptr;  // the library's own stream for this workspace (heads_ready steps)
  // two-kernel pipeline: scan on `stream`, post kernel on the caller's stream
  cudaEvent_t scanned[kCandSlots] = {};   // slot's candidate list is complete
  cudaEvent_t consumed[kCandSlots] = {};  // slot's post kernel is done (recorded on the caller's stream)
  bool consumed_valid[kCandSlots] = {};
  int covered = 0;  // scans to come whose slots are already known to be free
  // one-launch step
  cudaEvent_t done[kCandSlots] = {};  // the step kernel that used the slot has finished
  cudaEvent_t caller_mark = nullptr;
  int tiles_done_target[kCandSlots] = {};  // value of the slot's "tiles done" counter once its current scan is over
  int slot_free_target[kCandSlots] = {};   // value of the slot's "streams done" counter once its current readers are done
  int post_ticket_base[kCandSlots] = {};   // value of the slot's post-ticket counter before its current launch
  int seq = 0;            // step kernels launched so far
  int chain_streams = 0;  // batch size of the step kernel last enqueued on `stream` (0: none a new launch could overlap)
  int caller_seq = 0;     // caller-stream marks written so far
  void reset_counters() {
    for (int i = 0; i < kCandSlots; ++i) {
      tiles_done_target[i] = slot_free_target[i] = post_ticket_base[i] = 0;
      consumed_valid[i] = false;
    }
    seq = chain_streams = caller_seq = covered = 0;
  }
};
std::recursive_mutex& api_mutex();
size_t workspace_header_bytes(int num_streams);
int workspace_ctx(void* workspace, size_t workspace_bytes, int num_streams, cudaStream_t s, WorkspaceCtx** out);
int workspace_streams(WorkspaceCtx* c);  // creates the library's stream and events of the workspace if need be
int take_scan_slot(WorkspaceCtx* c, void* workspace, size_t workspace_bytes, int num_streams, int num_anchors, Workspace* ws);

// D1 + N1 only (defined in nms.cu): scans the head tensors and leaves the candidate list of every
// stream in `workspace`; *ws describes it for the NMS stage.  scan_stream != null: the scan goes to that
// stream (the library's own) instead of `stream`.
int launch_decode_stage(const void* head_p3, const void* head_p4, const void* head_p5, int head_dtype, int num_streams,
                        int img_h, int img_w, const rtm_nms_params* params, void* workspace, size_t workspace_bytes,
                        Workspace* ws, cudaStream_t stream, WorkspaceCtx** ctx = nullptr, cudaStream_t scan_stream = nullptr);

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

// Counting-rank sort of the n distinct keys[0..n) (layout below: class | score field | rank), the
// common case of the candidate sort.  Keys are dropped into kSortBuckets buckets by a monotone
// function of (class, score field) - so bucket order is key order -, bucket offsets come from one
// block scan, and every key finds its place inside its bucket by counting the smaller keys there.
// Eight barriers, whatever n is; the bitonic network above costs 15 to 28 and pads n to a power of
// two.  Returns false without touching keys when some bucket is crowded (many equal scores), so
// that the caller falls back to the network.  tmp: n keys; cnt: kSortBuckets + 1 ints.
constexpr int kSortBuckets = 2048;
constexpr int kSortBucketMax = 48;
template <int THREADS>
__device__ __forceinline__ bool bucket_sort_keys(uint64_t* keys, const int n, uint64_t* tmp, int* cnt, const uint32_t vmin,
                                                 const uint32_t vmax, const uint32_t* present, int* s_scan) {
  static_assert(kSortBuckets == 4 * THREADS, "one thread scans four buckets");
  constexpr int KMAX = kNmsSmemCand / THREADS;
  const int tid = threadIdx.x;
  // `present` (8 words, or null when the keys carry no class): bit c set = class c occurs; the
  // buckets are shared out evenly among the classes that occur
  uint32_t pw[8];
  int before[8], num_classes = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    pw[w] = present ? present[w] : 0u;
    before[w] = num_classes;
    num_classes += __popc(pw[w]);
  }
  if (!present) num_classes = 1;
  const int nbc = kSortBuckets / num_classes;  // score bins per class
  const float scale = static_cast<float>(nbc) / (static_cast<float>(vmax - vmin) + 1.f);
  for (int b = tid; b <= kSortBuckets; b += THREADS) cnt[b] = 0;
  __syncthreads();
  uint64_t mykey[KMAX];
  int bk[KMAX], ord[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int i = tid + k * THREADS;
    if (i < n) {
      const uint64_t key = keys[i];
      const uint32_t v = static_cast<uint32_t>(key >> 15);
      const int sb = min(nbc - 1, static_cast<int>(static_cast<float>(v - vmin) * scale));  // monotone in v
      mykey[k] = key;
      int cidx = 0;
      if (present) {
        const int cls = static_cast<int>(key >> 47);
#pragma unroll
        for (int w = 0; w < 8; ++w)
          if (w == (cls >> 5)) cidx = before[w] + __popc(pw[w] & ((1u << (cls & 31)) - 1u));
      }
      bk[k] = cidx * nbc + sb;
      ord[k] = atomicAdd(&cnt[bk[k]], 1);
    }
  }
  __syncthreads();
  const int c0 = cnt[4 * tid], c1 = cnt[4 * tid + 1], c2 = cnt[4 * tid + 2], c3 = cnt[4 * tid + 3];
  if (__syncthreads_or(max(max(c0, c1), max(c2, c3)) > kSortBucketMax)) return false;
  int tot;
  const int base = block_exclusive_sum<THREADS>(c0 + c1 + c2 + c3, s_scan, &tot);
  cnt[4 * tid] = base;
  cnt[4 * tid + 1] = base + c0;
  cnt[4 * tid + 2] = base + c0 + c1;
  cnt[4 * tid + 3] = base + c0 + c1 + c2;
  if (tid == THREADS - 1) cnt[kSortBuckets] = tot;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (tid + k * THREADS < n) tmp[cnt[bk[k]] + ord[k]] = mykey[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (tid + k * THREADS < n) {
      const int s = cnt[bk[k]], e = cnt[bk[k] + 1];
      int rank = 0;
      for (int j = s; j < e; ++j) rank += tmp[j] < mykey[k];
      keys[s + rank] = mykey[k];
    }
  }
  __syncthreads();
  return true;
}

// ---------------------------------------------------------------------------------------
// Greedy scan, class-parallel form.
//
// With the class offset (agnostic = 0) boxes of different classes cannot intersect as long as
// every coordinate stays inside (-kClassGuard, kClassGuard): consecutive offsets are kMaxWh
// apart.  Then torchvision's scan in global score order keeps exactly the union of what
// independent per-class scans keep, and its first max_det survivors are the max_det best-scored
// members of that union.  So: sort class-major (class, score desc, rank asc), let one warp
// resolve each class segment (suppression tests in registers, survivors broadcast by shuffle, no
// block barrier), hand segments too long for a warp to the block-wide scan, then order the
// survivors by (score, rank).  Every IoU is still formed on the offset boxes, so each decision
// is the one torchvision makes.  Streams that fail the guard, and agnostic NMS, take the
// block-wide scan in global score order.
// ---------------------------------------------------------------------------------------
constexpr float kClassGuard = 3800.f;
constexpr int kWarpSegMax = 512;  // class segments up to this long are resolved by a single warp
constexpr int kClsShift = 47;     // key = class (8 bits) | ~orderable(score) (32 bits) | rank (15 bits)
constexpr uint64_t kKeyNoClass = (1ull << kClsShift) - 1ull;

__device__ __forceinline__ float key_score(uint64_t key) {
  return float_from_orderable(~static_cast<uint32_t>(key >> kIdxBits));
}

// A class segment [s, e) of the sorted list is resolved in three steps (survivors = bits of kmask):
//   head    one warp settles the segment's first chunk (the positions that share s's mask word) -
//           its best-scored candidates, which is where nearly all survivors are;
//   filter  every thread tests its own candidate against the head survivors of its class: one
//           fully parallel pass that removes almost everything else;
//   tail    the warp walks what is left of its segment, chunk by chunk, exactly as the serial scan
//           would (against the survivors found after the head, then within the chunk).
// At most max_det survivors per class are recorded: later ones could never be reported.
__device__ __forceinline__ int warp_chunk_resolve(const float4 mb, const float ma, bool alive, int kc, const int max_det,
                                                  const float gate, uint32_t* keptbits) {
  const int lane = threadIdx.x & 31;
  uint32_t am = __ballot_sync(kFull, alive), kb_bits = 0u;
  while (am && kc < max_det) {
    const int l = __ffs(am) - 1;
    kb_bits |= 1u << l;
    ++kc;
    const float4 kb = make_float4(__shfl_sync(kFull, mb.x, l), __shfl_sync(kFull, mb.y, l), __shfl_sync(kFull, mb.z, l),
                                  __shfl_sync(kFull, mb.w, l));
    const float ka = __shfl_sync(kFull, ma, l);
    if (alive && lane > l && suppresses(kb, ka, mb, ma, gate)) alive = false;
    am = __ballot_sync(kFull, alive) & (l == 31 ? 0u : ~((2u << l) - 1u));
  }
  *keptbits = kb_bits;
  return kc;
}

// The head of segment [s, e) = its positions in the mask word of s plus, when that word is only
// partly the segment's, the following word: between 32 and 63 candidates (or the whole segment).
__device__ __forceinline__ int head_end(const int s) { return (s & ~31) + ((s & 31) ? 64 : 32); }

// head survivors recorded in mask word (s >> 5) + which, restricted to positions inside [s, e)
__device__ __forceinline__ uint32_t head_bits(const volatile uint32_t* kmask, const int s, const int e, const int which) {
  const int w = (s >> 5) + which;
  if (which == 1 && (!(s & 31) || (w << 5) >= e)) return 0u;
  uint32_t bits = kmask[w];
  if (which == 0) bits &= ~((1u << (s & 31)) - 1u);        // the word may hold the previous segment's tail
  if ((e >> 5) == w && (e & 31)) bits &= (1u << (e & 31)) - 1u;  // ... or the next segment's head
  return bits;
}

__device__ __forceinline__ void segment_head(const float4* sbox, const float* sarea, const int s, const int e,
                                             const int max_det, const float gate, uint32_t* kmask) {
  const int lane = threadIdx.x & 31;
  const int base = s & ~31;
  int i = base + lane;
  bool valid = i >= s && i < e;
  float4 mb = valid ? sbox[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float ma = valid ? sarea[i] : 0.f;
  uint32_t bits_a, bits_b;
  int kc = warp_chunk_resolve(mb, ma, valid, 0, max_det, gate, &bits_a);
  if (lane == 0 && bits_a) atomicOr(&kmask[base >> 5], bits_a);
  if ((s & 31) && base + 32 < e) {
    i += 32;
    valid = i < e;
    mb = valid ? sbox[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    ma = valid ? sarea[i] : 0.f;
    bool alive = valid;
    uint32_t bits = bits_a;
    while (bits) {
      const int p = base + __ffs(bits) - 1;
      bits &= bits - 1;
      if (alive && suppresses(sbox[p], sarea[p], mb, ma, gate)) alive = false;
    }
    warp_chunk_resolve(mb, ma, alive, kc, max_det, gate, &bits_b);
    if (lane == 0 && bits_b) atomicOr(&kmask[(base >> 5) + 1], bits_b);
  }
}

// candidate at sorted position i of segment [s, e), behind the head: does it survive the head?
__device__ __forceinline__ bool survives_head(const float4* sbox, const float* sarea, const int i, const int s, const int e,
                                              const float gate, const uint32_t* kmask) {
  const float4 mb = sbox[i];
  const float ma = sarea[i];
  bool alive = true;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    uint32_t bits = head_bits(kmask, s, e, which);
    while (bits) {
      const int p = (((s >> 5) + which) << 5) + __ffs(bits) - 1;
      bits &= bits - 1;
      if (suppresses(sbox[p], sarea[p], mb, ma, gate)) alive = false;
    }
  }
  return alive;
}

__device__ __forceinline__ void segment_tail(const float4* sbox, const float* sarea, const int s, const int e,
                                             const int max_det, const float gate, uint32_t* kmask, uint32_t* alive0) {
  const int lane = threadIdx.x & 31;
  const volatile uint32_t* vmask = kmask;
  const int first = head_end(s);  // first chunk behind the head (word-aligned)
  int kc = __popc(head_bits(vmask, s, e, 0)) + __popc(head_bits(vmask, s, e, 1));
  for (int base = first; base < e; base += 32) {
    const uint32_t word = alive0[base >> 5];
    const int i = base + lane;
    const bool valid = i < e && ((word >> lane) & 1u);
    __syncwarp();
    if (lane == 0) alive0[base >> 5] = 0u;  // consumed (bits past e belong to the next segment's head: already zero)
    if (!__any_sync(kFull, valid) || kc >= max_det) continue;
#ifdef RTM_TIMELINE
    if (lane == 0 && g_timeline) {
      atomicAdd(&g_timeline[blockIdx.x * 32 + 22], 1ull);                                         // chunks with work
      atomicAdd(&g_timeline[blockIdx.x * 32 + 23], (unsigned long long)__popc(word));  // alive candidates in it
    }
    const unsigned long long t_a = clock64();
#endif
    const float4 mb = valid ? sbox[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ma = valid ? sarea[i] : 0.f;
    bool alive = valid;
    for (int w = first >> 5; w < (base >> 5); ++w) {  // survivors found behind the head
      uint32_t bits = vmask[w];
      while (bits) {
        const int p = (w << 5) + __ffs(bits) - 1;
        bits &= bits - 1;
        if (alive && suppresses(sbox[p], sarea[p], mb, ma, gate)) alive = false;
      }
    }
#ifdef RTM_TIMELINE
    const unsigned long long t_b = clock64();
#endif
    uint32_t keptbits;
    kc = warp_chunk_resolve(mb, ma, alive, kc, max_det, gate, &keptbits);
    if (lane == 0 && keptbits) atomicOr(&kmask[base >> 5], keptbits);
    __syncwarp();
#ifdef RTM_TIMELINE
    if (lane == 0 && g_timeline) {
      atomicAdd(&g_timeline[blockIdx.x * 32 + 24], (unsigned long long)__popc(keptbits));   // survivors found in the tail
      atomicAdd(&g_timeline[blockIdx.x * 32 + 25], t_b - t_a);                               // cycles: vs earlier tail survivors
      atomicAdd(&g_timeline[blockIdx.x * 32 + 26], (unsigned long long)clock64() - t_b);     // cycles: chunk resolve
    }
#endif
  }
#ifdef RTM_TIMELINE
  if (lane == 0 && g_timeline) atomicMax(&g_timeline[blockIdx.x * 32 + 27], (unsigned long long)(e - s));  // longest segment
#endif
}

// Block-wide greedy scan over the candidates whose bit is set in alive0 (register-resident:
// thread tid owns sorted positions tid + k*THREADS).  One survivor per barrier round, the alive
// bitmask rebuilt by warp ballots into a ping-pong buffer.  TO_MASK: survivors are recorded as
// bits of kmask (no limit); otherwise in s_keep[0..max_det) in scan order.  Returns the number
// of survivors.  alive0 / alive1: words entries each.
template <int THREADS, bool TO_MASK>
__device__ __forceinline__ int block_scan_nms(const float4* sbox, const float* sarea, const int n, const int words,
                                              uint32_t* alive0, uint32_t* alive1, const int max_det, const float gate,
                                              int* s_keep, uint32_t* kmask) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = THREADS / 32;
  constexpr int KMAX = kNmsSmemCand / THREADS;
  float4 mybox[KMAX];
  float myarea[KMAX];
  bool myalive[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = tid + k * THREADS;
    myalive[k] = j < n && ((alive0[j >> 5] >> (j & 31)) & 1u);
    mybox[k] = myalive[k] ? sbox[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    myarea[k] = myalive[k] ? sarea[j] : 0.f;
  }
  const int kmax = (n + THREADS - 1) / THREADS;  // slots that can hold a candidate at all
  uint32_t* cur = alive0;
  uint32_t* nxt = alive1;
  int kept = 0, w0 = 0;
  while (TO_MASK || kept < max_det) {
    while (w0 < words && cur[w0] == 0u) ++w0;
    if (w0 >= words) break;
    const int c = (w0 << 5) + __ffs(cur[w0]) - 1;
    const float4 kb = sbox[c];
    const float ka = sarea[c];
    if (tid == 0) {
      if (TO_MASK) kmask[c >> 5] |= 1u << (c & 31);
      else s_keep[kept] = c;
    }
    ++kept;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < kmax) {
        const int j = tid + k * THREADS;
        bool alive = myalive[k];
        if (alive) alive = j != c && !suppresses(kb, ka, mybox[k], myarea[k], gate);
        myalive[k] = alive;
        const uint32_t nw = __ballot_sync(kFull, alive);
        if (lane == 0 && k * kWarps + warp < words) nxt[k * kWarps + warp] = nw;
      }
    }
    __syncthreads();
    uint32_t* t = cur;
    cur = nxt;
    nxt = t;
  }
  return kept;
}

// The NMS proper for a stream with n > 0 candidates whose ranks are already expanded into
// loc[0..n).  IN_SMEM: working arrays live in `smem` (kNmsSmemBytes, n <= kNmsSmemCand);
// otherwise in the workspace spill arrays.
template <int THREADS, bool IN_SMEM>
__device__ __forceinline__ int nms_run(const Workspace& ws, const rtm_nms_params& prm, const float iou_gate,
                                       const NmsOut& out, const int b, const int n, unsigned char* smem,
                                       int* s_keep, int* s_scan) {
  static_assert(kNmsSmemCand % THREADS == 0 && kNmsSmemCand / 32 <= 2048, "per-thread candidate slots are kNmsSmemCand / THREADS");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = THREADS / 32;
  const int A = ws.num_anchors, W = ws.words;
  const size_t a0 = static_cast<size_t>(b) * A;
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
  // class tables of the class-parallel scan (IN_SMEM only)
  __shared__ int c_start[256], c_end[256], c_list[256];
  __shared__ int c_count, c_large;
  __shared__ uint32_t kmask[kNmsSmemCand / 32];
  __shared__ uint32_t s_vmin, s_vmax;  // range of the score field of the keys
  __shared__ uint32_t s_present[8];    // classes that occur among the candidates

  // ---- stage candidates (one global round trip); decide whether classes are independent ----
  int P = 32;
  while (P < n) P <<= 1;
  bool guard_ok = true;
  uint32_t vlo = 0xffffffffu, vhi = 0u;
  if (tid == 0) {
    s_vmin = 0xffffffffu;
    s_vmax = 0u;
  }
  if (tid < 8) s_present[tid] = 0u;
  __syncthreads();
  uint32_t my_score[IN_SMEM ? kNmsSmemCand / THREADS : 1];
  for (int i = tid, k = 0; i < P; i += THREADS, ++k) {
    if (i < n) {
      const int an = loc[i];
      const float sc = ws.score[a0 + an];
      const float4 bx = ws.box[a0 + an];
      ubox[i] = bx;
      const int cls = ws.cls[a0 + an];
      loc[i] = an | (cls << 16);
      if (IN_SMEM) atomicOr(&s_present[(cls >> 5) & 7], 1u << (cls & 31));
      guard_ok = guard_ok && fabsf(bx.x) < kClassGuard && fabsf(bx.y) < kClassGuard && fabsf(bx.z) < kClassGuard &&
                 fabsf(bx.w) < kClassGuard;  // NaN fails
      if (IN_SMEM) {
        my_score[k] = ~float_orderable(sc);
        vlo = min(vlo, my_score[k]);
        vhi = max(vhi, my_score[k]);
      } else keys[i] = (static_cast<uint64_t>(~float_orderable(sc)) << kIdxBits) | static_cast<uint32_t>(i);
    } else if (!IN_SMEM) {
      keys[i] = ~0ull;
    }
  }
  bool by_class = false;
  if (IN_SMEM) {
    vlo = __reduce_min_sync(kFull, vlo);
    vhi = __reduce_max_sync(kFull, vhi);
    if (lane == 0) {
      atomicMin(&s_vmin, vlo);
      atomicMax(&s_vmax, vhi);
    }
    by_class = __syncthreads_and(guard_ok) && !prm.agnostic;
    // sort keys: class-major when classes are independent, else global score order; ties by
    // ascending rank (= torchvision's stable sort of the filtered list)
    for (int i = tid, k = 0; i < P; i += THREADS, ++k) {
      uint64_t key = ~0ull;
      if (i < n) {
        key = (static_cast<uint64_t>(my_score[k]) << kIdxBits) | static_cast<uint32_t>(i);
        if (by_class) key |= static_cast<uint64_t>(loc[i] >> 16) << kClsShift;
      }
      keys[i] = key;
    }
    if (tid == 0) {
      c_count = 0;
      c_large = 0;
    }
  }
  __syncthreads();
  RTM_TL(3);
  if (IN_SMEM) {
    // sbox is not needed before the gather below: its first half holds the scattered keys, the
    // bucket offsets follow
    uint64_t* tmp = reinterpret_cast<uint64_t*>(sbox);
    int* cnt = reinterpret_cast<int*>(tmp + kNmsSmemCand);
    if (!bucket_sort_keys<THREADS>(keys, n, tmp, cnt, s_vmin, s_vmax, by_class ? s_present : nullptr, s_scan))
      bitonic_sort_keys<THREADS>(keys, P);
  } else {
    bitonic_sort_keys<THREADS>(keys, P);
  }
  RTM_TL(4);

  // ---- boxes in sorted order with the class offset added in float32, areas as torchvision ----
  const int words = (n + 31) >> 5;
  for (int i = tid; i < n; i += THREADS) {
    const uint64_t key = keys[i];
    const int r = static_cast<int>(key & ((1u << kIdxBits) - 1u));
    float4 bx = ubox[r];
    const int cls = loc[r] >> 16;
    const float off = prm.agnostic ? 0.f : __fmul_rn(static_cast<float>(cls), kMaxWh);
    bx.x = __fadd_rn(bx.x, off);
    bx.y = __fadd_rn(bx.y, off);
    bx.z = __fadd_rn(bx.z, off);
    bx.w = __fadd_rn(bx.w, off);
    sbox[i] = bx;
    sarea[i] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    if (IN_SMEM && by_class) {  // segment boundaries of the class-major order
      const int prev = i > 0 ? static_cast<int>(keys[i - 1] >> kClsShift) : -1;
      const int next = i + 1 < n ? static_cast<int>(keys[i + 1] >> kClsShift) : -1;
      if (prev != cls) {
        c_start[cls] = i;
        c_list[atomicAdd(&c_count, 1)] = cls;
      }
      if (next != cls) c_end[cls] = i + 1;
    }
  }
  for (int w = tid; w < words; w += THREADS) {
    const int rem = n - (w << 5);
    alive0[w] = (IN_SMEM && by_class) ? 0u : (rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u));
    if (IN_SMEM) kmask[w] = 0u;
  }
  __syncthreads();
  RTM_TL(5);

  int kept = 0;
  if (IN_SMEM && by_class) {
    // ---- head: one warp per class segment ----
    const int nseg = c_count;
    for (int k = warp; k < nseg; k += kWarps) {
      const int cls = c_list[k];
      segment_head(sbox, sarea, c_start[cls], c_end[cls], max_det, iou_gate, kmask);
    }
    __syncthreads();
    // ---- filter: every candidate behind its segment's head chunk against the head survivors ----
    for (int i0 = 0; i0 < n; i0 += THREADS) {
      const int i = i0 + tid;
      bool alive = false;
      if (i < n) {
        const int cls = static_cast<int>(keys[i] >> kClsShift), s = c_start[cls];
        if (i >= head_end(s)) alive = survives_head(sbox, sarea, i, s, c_end[cls], iou_gate, kmask);
      }
      const uint32_t word = __ballot_sync(kFull, alive);
      if (lane == 0 && (i0 + tid) < ((n + 31) & ~31)) alive0[(i0 + tid) >> 5] = word;
    }
    __syncthreads();
    RTM_TL(7);
    // ---- tail: what is left of each segment; long segments are left to the block-wide scan ----
    for (int k = warp; k < nseg; k += kWarps) {
      const int cls = c_list[k], s = c_start[cls], e = c_end[cls];
      if (e - s <= kWarpSegMax) segment_tail(sbox, sarea, s, e, max_det, iou_gate, kmask, alive0);
      else if (lane == 0) c_large = 1;
    }
    __syncthreads();
    if (c_large) block_scan_nms<THREADS, true>(sbox, sarea, n, words, alive0, alive1, max_det, iou_gate, s_keep, kmask);
    __syncthreads();
    RTM_TL(8);
    // ---- survivors of all classes -> (score desc, rank asc) order, first max_det ----
    int* kl = reinterpret_cast<int*>(sarea);               // sorted positions of the survivors (sbox / sarea are dead now)
    uint64_t* kk = reinterpret_cast<uint64_t*>(sbox);      // their keys without the class field
    int K = 0;
    for (int w0 = 0; w0 < words; w0 += THREADS) {
      const int w = w0 + tid;
      uint32_t m = w < words ? kmask[w] : 0u;
      int tot;
      int r = K + block_exclusive_sum<THREADS>(__popc(m), s_scan, &tot);
      while (m) {
        const int pos = (w << 5) + __ffs(m) - 1;
        m &= m - 1;
        kl[r] = pos;
        kk[r] = keys[pos] & kKeyNoClass;
        ++r;
      }
      K += tot;
    }
    __syncthreads();
    for (int i = tid; i < K; i += THREADS) {
      const uint64_t mine = kk[i];
      int rank = 0;
      for (int j = 0; j < K; ++j) rank += kk[j] < mine;  // keys are distinct (the rank field)
      if (rank < max_det) s_keep[rank] = kl[i];
    }
    kept = min(K, max_det);
  } else if (IN_SMEM) {
    kept = block_scan_nms<THREADS, false>(sbox, sarea, n, words, alive0, alive1, max_det, iou_gate, s_keep, kmask);
  } else {
    // ---- greedy scan on the spill arrays, one survivor per iteration; alive bits rebuilt by ballot ----
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
  RTM_TL(6);

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
    out.conf[o0 + o] = key_score(key);
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
struct NoPrologue {
  __device__ __forceinline__ void operator()() const {}
};

// `prologue` runs right after the stream's mask words have been requested and before anything waits
// for them: the fused kernel hangs its table prefetches there, so that their load chains and the
// mask load travel together.
template <int THREADS, typename Prologue = NoPrologue>
__device__ __forceinline__ int nms_stream(const Workspace& ws, const rtm_nms_params& prm, const float iou_gate,
                                          const NmsOut& out, const int b, unsigned char* smem, int* s_keep,
                                          int* s_scan, Prologue prologue = Prologue()) {
  const int tid = threadIdx.x;
  const int A = ws.num_anchors, W = ws.words;
  const uint32_t* mask = ws.mask + static_cast<size_t>(b) * W;
  if (ws.tile_counter && b == 0 && tid == 0) *ws.tile_counter = 0;  // the scan that filled this list is over: re-arm its tickets
  // ---- one pass over the stream's candidate mask: words stay in registers, ranks by block scan ----
  constexpr int kIters = (kMaxAnchors / 32 + THREADS - 1) / THREADS;
  uint32_t mw[kIters];
  int rb[kIters];
  int n = 0;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int w = it * THREADS + tid;
    mw[it] = w < W ? mask[w] : 0u;
  }
  prologue();
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    rb[it] = 0;
    if (it * THREADS < W) {  // block-uniform
      const int w = it * THREADS + tid;
      uint32_t m = mw[it];
      const int valid = A - (w << 5);
      if (w < W && valid < 32) m &= (1u << valid) - 1u;
      int tot;
      rb[it] = n + block_exclusive_sum<THREADS>(__popc(m), s_scan, &tot);
      mw[it] = m;
      n += tot;
    }
  }
  RTM_TL(1);
  if (n == 0) {
    if (tid == 0) out.count[b] = 0;
    __syncthreads();
    return 0;
  }
  const bool in_smem = n <= kNmsSmemCand;
  int32_t* loc = in_smem ? reinterpret_cast<int32_t*>(smem + (sizeof(uint64_t) + 2 * sizeof(float4) + sizeof(float)) * kNmsSmemCand)
                         : ws.loc + static_cast<size_t>(b) * A;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    uint32_t m = mw[it];
    int r = rb[it];
    const int w = it * THREADS + tid;
    while (m) {
      loc[r++] = (w << 5) + __ffs(m) - 1;
      m &= m - 1;
    }
  }
  __syncthreads();
  RTM_TL(2);
  if (in_smem) return nms_run<THREADS, true>(ws, prm, iou_gate, out, b, n, smem, s_keep, s_scan);
  return nms_run<THREADS, false>(ws, prm, iou_gate, out, b, n, smem, s_keep, s_scan);
}

}  // namespace rtm
