// T1..T4 for one stream as a device function (used by track_step_kernel and by the fused
// post-backbone kernel).  See track.cu for the semantics and the reference citations.
#pragma once

#include <limits.h>

#include "rtm_common.cuh"

namespace rtm {

constexpr int kStage2Flag = 0x40000000;

struct TrackArgs {
  rtm_track_table tin, tout;
  const float* det_xyxy;
  const float* det_conf;
  const int32_t* det_cls;
  const int32_t* det_count;
  int32_t det_stride;
  float track_thresh, match_thresh;
  int32_t track_buffer;
  int32_t* det_track_id;
  int32_t* det_kind;
  int32_t* src_row;
  int32_t* status;
};

// tracker.py:153-161 on one pair.  Non-overlapping pairs give exactly +0 (finite boxes with
// non-negative area), so the division is skipped for them.
__device__ __forceinline__ float pair_iou(const float4 a, const float area_a, const float4 b,
                                          const float area_b) {
  const float iw = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float ih = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(iw, ih);
  if (!(inter > 0.f)) return 0.f;
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  return __fdiv_rn(inter, __fadd_rn(uni, 1e-6f));
}

__device__ __forceinline__ float box_area(const float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// One association stage: rows = tracks with s_match[t] < 0, columns = s_list[0..m).
// On return s_match[t] holds (det index | flag) for the rows that won their column.
template <int THREADS>
__device__ __forceinline__ void associate(const float4* __restrict__ trk_box, int T,
                                          const float4* s_box, const float* s_area,
                                          const int* s_list, int m, int* s_win, int* s_match,
                                          float thresh, int flag) {
  const int tid = threadIdx.x;
  for (int j = tid; j < m; j += THREADS) s_win[j] = INT_MAX;
  __syncthreads();
  for (int t = tid; t < T; t += THREADS) {
    if (s_match[t] >= 0) continue;
    const float4 a = trk_box[t];
    const float area_a = box_area(a);
    int d0 = s_list[0];
    float best = pair_iou(a, area_a, s_box[d0], s_area[d0]);
    int bj = 0;
    for (int j = 1; j < m; ++j) {
      const int d = s_list[j];
      const float v = pair_iou(a, area_a, s_box[d], s_area[d]);
      if (v > best) {  // strict: first arg-max, np.argmax semantics (tracker.py:187)
        best = v;
        bj = j;
      }
    }
    if (best >= thresh) {  // tracker.py:188, float32 compare
      atomicMin(&s_win[bj], t);
      s_match[t] = -2 - bj;  // bidding for column bj
    }
  }
  __syncthreads();
  for (int t = tid; t < T; t += THREADS) {
    const int v = s_match[t];
    if (v <= -2) {
      const int bj = -2 - v;
      s_match[t] = (s_win[bj] == t) ? (s_list[bj] | flag) : -1;
    }
  }
  __syncthreads();
}

// One stream.  `smem_raw`: track_smem_bytes(det_stride, capacity) bytes of shared memory,
// 16-byte aligned.  All THREADS threads of the block must call it (block-uniform control flow).
template <int THREADS>
__device__ __forceinline__ void track_stream(const TrackArgs& a, const int b, unsigned char* smem_raw) {
  const int tid = threadIdx.x;
  const int cap = a.tin.capacity, S = a.det_stride;

  float4* s_box = reinterpret_cast<float4*>(smem_raw);  // S
  float* s_area = reinterpret_cast<float*>(s_box + S);  // S
  int* s_hi = reinterpret_cast<int*>(s_area + S);       // S  det index of the j-th high det
  int* s_lo = s_hi + S;                                 // S  ... low det; later: birth list
  int* s_win = s_lo + S;                                // S  column winners of a stage
  int* s_born = s_win + S;                              // S  1 if high column j is unmatched
  int* s_match = s_born + S;                            // cap
  int* s_scan = s_match + cap;                          // 33

  const size_t row0 = static_cast<size_t>(b) * cap;
  const size_t det0 = static_cast<size_t>(b) * S;
  const float4* in_box = reinterpret_cast<const float4*>(a.tin.xyxy) + row0;
  float4* out_box = reinterpret_cast<float4*>(a.tout.xyxy) + row0;
  const float4* det_box = reinterpret_cast<const float4*>(a.det_xyxy) + det0;

  const int T = min(a.tin.count[b], cap);
  const int next_id = a.tin.next_id[b];
  int n = a.det_count[b];
  int st = 0;
  if (n > S) {
    n = S;
    st |= RTM_STATUS_DET_OVERFLOW;
  }
  if (n < 0) n = 0;

  if (n == 0) {
    // tracker.py:70-73: age only, nothing is pruned on an empty frame
    for (int t = tid; t < T; t += THREADS) {
      a.tout.track_id[row0 + t] = a.tin.track_id[row0 + t];
      out_box[t] = in_box[t];
      a.tout.confidence[row0 + t] = a.tin.confidence[row0 + t];
      a.tout.class_id[row0 + t] = a.tin.class_id[row0 + t];
      a.tout.age[row0 + t] = a.tin.age[row0 + t];
      a.tout.time_since_update[row0 + t] = a.tin.time_since_update[row0 + t] + 1;
      if (a.src_row) a.src_row[row0 + t] = t;
    }
    if (tid == 0) {
      a.tout.count[b] = T;
      a.tout.next_id[b] = next_id;
      if (st && a.status) atomicOr(&a.status[b], st);
    }
    return;
  }

  // ---- T1: stage detections, split high / low in order (tracker.py:76-85) ---------------
  int H = 0, L = 0;
  for (int r0 = 0; r0 < n; r0 += THREADS) {
    const int d = r0 + tid;
    const bool valid = d < n;
    bool hi = false;
    if (valid) {
      const float4 bx = det_box[d];
      s_box[d] = bx;
      s_area[d] = box_area(bx);
      hi = a.det_conf[det0 + d] >= a.track_thresh;
      if (a.det_track_id) a.det_track_id[det0 + d] = 0;
      if (a.det_kind) a.det_kind[det0 + d] = RTM_DET_NONE;
    }
    int tot;
    const int p = block_exclusive_count(hi, s_scan, &tot);
    if (hi) s_hi[H + p] = d;
    if (valid && !hi) s_lo[L + (tid - p)] = d;
    const int in_round = min(THREADS, n - r0);
    H += tot;
    L += in_round - tot;
  }
  for (int t = tid; t < T; t += THREADS) s_match[t] = -1;
  for (int j = tid; j < H; j += THREADS) s_born[j] = 1;
  __syncthreads();

  // ---- stage 1: all retained tracks x high detections (tracker.py:91-104) ---------------
  if (T > 0 && H > 0) {
    associate<THREADS>(in_box, T, s_box, s_area, s_hi, H, s_win, s_match, a.match_thresh, 0);
    for (int j = tid; j < H; j += THREADS) s_born[j] = (s_win[j] == INT_MAX);
    __syncthreads();
  }
  // ---- stage 2: still-unmatched tracks x low detections, same threshold (tracker.py:109-123)
  if (T > 0 && L > 0) {
    associate<THREADS>(in_box, T, s_box, s_area, s_lo, L, s_win, s_match, a.match_thresh,
                       kStage2Flag);
  }

  // ---- births: unmatched high detections in ascending order (tracker.py:126-135) --------
  int NB = 0;
  int* s_birth = s_lo;  // low list is dead from here on
  for (int r0 = 0; r0 < H; r0 += THREADS) {
    const int j = r0 + tid;
    const bool born = j < H && s_born[j];
    int tot;
    const int p = block_exclusive_count(born, s_scan, &tot);
    if (born) s_birth[NB + p] = s_hi[j];
    NB += tot;
  }
  __syncthreads();

  // ---- update, age, prune, compact (tracker.py:99-104, 138-139, 144-147) ---------------
  const bool birth_survives = 1 <= a.track_buffer;
  int kept = 0;
  const int V = T + NB;
  for (int r0 = 0; r0 < V; r0 += THREADS) {
    const int v = r0 + tid;
    bool keep = false;
    int id = 0, cls = 0, age = 0, tsu = 0, det = -1, kind = RTM_DET_NONE;
    float conf = 0.f;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < T) {
      id = a.tin.track_id[row0 + v];
      const int m = s_match[v];
      if (m >= 0) {
        det = m & ~kStage2Flag;
        kind = (m & kStage2Flag) ? RTM_DET_STAGE2 : RTM_DET_STAGE1;
        box = s_box[det];
        conf = a.det_conf[det0 + det];
        cls = a.det_cls[det0 + det];
        age = a.tin.age[row0 + v] + 1;
        tsu = 1;
      } else {
        box = in_box[v];
        conf = a.tin.confidence[row0 + v];
        cls = a.tin.class_id[row0 + v];
        age = a.tin.age[row0 + v];
        tsu = a.tin.time_since_update[row0 + v] + 1;
      }
      keep = tsu <= a.track_buffer;
    } else if (v < V) {
      det = s_birth[v - T];
      kind = RTM_DET_BIRTH;
      id = next_id + (v - T);
      box = s_box[det];
      conf = a.det_conf[det0 + det];
      cls = a.det_cls[det0 + det];
      age = 1;
      tsu = 1;
      keep = birth_survives;
    }
    if (det >= 0) {
      if (a.det_track_id) a.det_track_id[det0 + det] = id;
      if (a.det_kind) a.det_kind[det0 + det] = kind;
    }
    int tot;
    const int p = kept + block_exclusive_count(keep, s_scan, &tot);
    if (keep && p < cap) {
      a.tout.track_id[row0 + p] = id;
      out_box[p] = box;
      a.tout.confidence[row0 + p] = conf;
      a.tout.class_id[row0 + p] = cls;
      a.tout.age[row0 + p] = age;
      a.tout.time_since_update[row0 + p] = tsu;
      if (a.src_row) a.src_row[row0 + p] = v < T ? v : -1;
    }
    kept += tot;
  }
  if (tid == 0) {
    if (kept > cap) {
      st |= RTM_STATUS_TRACK_OVERFLOW;
      kept = cap;
    }
    a.tout.count[b] = kept;
    a.tout.next_id[b] = next_id + NB;
    if (st && a.status) atomicOr(&a.status[b], st);
  }
}

inline size_t track_smem_bytes(int det_stride, int capacity) {
  return static_cast<size_t>(det_stride) * (16 + 4 + 4 * 4 + 4) + static_cast<size_t>(capacity) * 4 + 40 * 4;
}

}  // namespace rtm
