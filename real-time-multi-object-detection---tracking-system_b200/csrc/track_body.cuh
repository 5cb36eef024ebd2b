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
  // opt-in motion model (row K): all four null = the reference's behaviour
  const float* kf_mean_in;
  const float* kf_cov_in;
  float* kf_mean_out;
  float* kf_cov_out;
  int32_t assignment;  // RTM_ASSIGN_*
  double cost_limit;   // RTM_ASSIGN_OPTIMAL: lap's cost_limit = 1 - match_thresh (tracker.py:170), in double
  // RTM_ASSIGN_OPTIMAL: global scratch for stages beyond the shared-memory solver (null: they set RTM_STATUS_ASSIGN_LIMIT)
  unsigned char* assign_scratch;
  size_t assign_scratch_per_stream;
};

// ---- constant-velocity filter of ByteTrack (xyah), stored as four (position, velocity) filters.
// The arithmetic below is restated operation for operation, in float32, by oracle/kalman_ref.py
// (decoupled form) and checked there against the canonical 8 x 8 matrix form.
struct KalmanTrack {
  float m[8];   // x, y, a, h, vx, vy, va, vh
  float c[12];  // per coordinate: var(pos), cov(pos, vel), var(vel)
};

__device__ __forceinline__ void kalman_load(KalmanTrack& k, const float* mean, const float* cov, size_t row) {
  const float4* m4 = reinterpret_cast<const float4*>(mean + row * 8);
  const float4* c4 = reinterpret_cast<const float4*>(cov + row * 12);
  const float4 a = m4[0], b = m4[1], c0 = c4[0], c1 = c4[1], c2 = c4[2];
  k.m[0] = a.x; k.m[1] = a.y; k.m[2] = a.z; k.m[3] = a.w;
  k.m[4] = b.x; k.m[5] = b.y; k.m[6] = b.z; k.m[7] = b.w;
  k.c[0] = c0.x; k.c[1] = c0.y; k.c[2] = c0.z; k.c[3] = c0.w;
  k.c[4] = c1.x; k.c[5] = c1.y; k.c[6] = c1.z; k.c[7] = c1.w;
  k.c[8] = c2.x; k.c[9] = c2.y; k.c[10] = c2.z; k.c[11] = c2.w;
}

__device__ __forceinline__ void kalman_store(const KalmanTrack& k, float* mean, float* cov, size_t row) {
  float4* m4 = reinterpret_cast<float4*>(mean + row * 8);
  float4* c4 = reinterpret_cast<float4*>(cov + row * 12);
  m4[0] = make_float4(k.m[0], k.m[1], k.m[2], k.m[3]);
  m4[1] = make_float4(k.m[4], k.m[5], k.m[6], k.m[7]);
  c4[0] = make_float4(k.c[0], k.c[1], k.c[2], k.c[3]);
  c4[1] = make_float4(k.c[4], k.c[5], k.c[6], k.c[7]);
  c4[2] = make_float4(k.c[8], k.c[9], k.c[10], k.c[11]);
}

constexpr float kStdPos = 1.f / 20.f, kStdVel = 1.f / 160.f;

__device__ __forceinline__ void box_to_xyah(const float4 b, float (&z)[4]) {
  const float w = b.z - b.x, h = b.w - b.y;
  z[0] = b.x + w * 0.5f;
  z[1] = b.y + h * 0.5f;
  z[2] = w / h;
  z[3] = h;
}

__device__ __forceinline__ float4 xyah_to_box(const float cx, const float cy, const float a, const float h) {
  const float w = a * h;
  const float x1 = cx - w * 0.5f, y1 = cy - h * 0.5f;
  return make_float4(x1, y1, x1 + w, y1 + h);
}

// KalmanFilter.initiate
__device__ __forceinline__ void kalman_initiate(KalmanTrack& k, const float4 box) {
  float z[4];
  box_to_xyah(box, z);
  const float sp = (2.f * kStdPos) * z[3], sv = (10.f * kStdVel) * z[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    k.m[i] = z[i];
    k.m[4 + i] = 0.f;
    k.c[3 * i] = i == 2 ? 1e-2f * 1e-2f : sp * sp;
    k.c[3 * i + 1] = 0.f;
    k.c[3 * i + 2] = i == 2 ? 1e-5f * 1e-5f : sv * sv;
  }
}

// STrack.predict + KalmanFilter.predict: a track that missed the previous frame has vh zeroed
__device__ __forceinline__ void kalman_predict(KalmanTrack& k, const int tsu_in) {
  if (tsu_in > 1) k.m[7] = 0.f;
  const float sp = kStdPos * k.m[3], sv = kStdVel * k.m[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float qp = i == 2 ? 1e-2f * 1e-2f : sp * sp, qv = i == 2 ? 1e-5f * 1e-5f : sv * sv;
    const float pp = k.c[3 * i], pv = k.c[3 * i + 1], vv = k.c[3 * i + 2];
    k.m[i] = k.m[i] + k.m[4 + i];
    k.c[3 * i] = ((pp + pv) + (pv + vv)) + qp;
    k.c[3 * i + 1] = pv + vv;
    k.c[3 * i + 2] = vv + qv;
  }
}

// KalmanFilter.project + update with the measurement box
__device__ __forceinline__ void kalman_update(KalmanTrack& k, const float4 box) {
  float z[4];
  box_to_xyah(box, z);
  const float sr = kStdPos * k.m[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float r = i == 2 ? 1e-1f * 1e-1f : sr * sr;
    const float pp = k.c[3 * i], pv = k.c[3 * i + 1], vv = k.c[3 * i + 2];
    const float s = pp + r;
    const float kp = pp / s, kv = pv / s;
    const float y = z[i] - k.m[i];
    k.m[i] = k.m[i] + kp * y;
    k.m[4 + i] = k.m[4 + i] + kv * y;
    k.c[3 * i] = pp - kp * pp;
    k.c[3 * i + 1] = pv - kp * pv;
    k.c[3 * i + 2] = vv - kv * pv;
  }
}

// box the association sees for a row: predicted from the filter state (mean only)
__device__ __forceinline__ float4 kalman_predicted_box(const float* mean, size_t row, const int tsu_in) {
  const float4* m4 = reinterpret_cast<const float4*>(mean + row * 8);
  const float4 p = m4[0], v = m4[1];
  const float vh = tsu_in > 1 ? 0.f : v.w;
  return xyah_to_box(p.x + v.x, p.y + v.y, p.z + v.z, p.w + vh);
}

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

// The head of a stream's input table, staged in shared memory ahead of use: the fused kernel
// issues these loads before the NMS stage so that their latency is off the critical path.
#ifndef RTM_TRACK_PREF_ROWS
#define RTM_TRACK_PREF_ROWS 256
#endif
constexpr int kTrackPrefRows = RTM_TRACK_PREF_ROWS;
struct TrackPrefetch {
  float4 box[kTrackPrefRows];   // stored box (tracker.py:100: the last matched detection)
  float4 abox[kTrackPrefRows];  // box the association sees: the stored one, or the filter's prediction
  int32_t track_id[kTrackPrefRows];
  float confidence[kTrackPrefRows];
  int32_t class_id[kTrackPrefRows];
  int32_t age[kTrackPrefRows];
  int32_t tsu[kTrackPrefRows];
  int32_t count, next_id;
};

// All threads call it; the caller provides a block barrier before track_stream reads `pf`.
template <int THREADS>
__device__ __forceinline__ void track_prefetch(const TrackArgs& a, const int b, TrackPrefetch* pf) {
  const rtm_track_table& tin = a.tin;
  const int tid = threadIdx.x;
  const size_t row0 = static_cast<size_t>(b) * tin.capacity;
  const int T = min(tin.count[b], tin.capacity);
  const float4* in_box = reinterpret_cast<const float4*>(tin.xyxy) + row0;
  for (int t = tid; t < min(T, kTrackPrefRows); t += THREADS) {
    const int tsu = tin.time_since_update[row0 + t];
    const float4 stored = in_box[t];
    pf->box[t] = stored;
    pf->abox[t] = a.kf_mean_in ? kalman_predicted_box(a.kf_mean_in, row0 + t, tsu) : stored;
    pf->track_id[t] = tin.track_id[row0 + t];
    pf->confidence[t] = tin.confidence[row0 + t];
    pf->class_id[t] = tin.class_id[row0 + t];
    pf->age[t] = tin.age[row0 + t];
    pf->tsu[t] = tsu;
  }
  if (tid == 0) {
    pf->count = T;
    pf->next_id = tin.next_id[b];
  }
}

// Columns of a stage binned by the x coordinate of their box centre, for stages with many columns (a crowd):
// a row then only looks at the columns whose box can intersect its own - every other pair has IoU exactly +0 and
// cannot be the row's first arg-max unless the whole row is zero, in which case the row matches nothing anyway
// (match_thresh > 0).  Exact by construction: a column d intersects row t only if
//   t.x1 - w_d / 2 < cx_d < t.x2 + w_d / 2,  w_d <= w_max,
// so the bins covering [t.x1 - w_max / 2, t.x2 + w_max / 2], widened by one bin on each side against rounding, hold
// every such column.  The order inside a bin is arbitrary (atomics); ties are resolved by the column's index in the
// stage's list, as np.argmax does.
constexpr int kColBins = 64;
constexpr int kBinMinCols = 128;  // stages with fewer columns scan them all
struct ColBins {
  int start[kColBins + 1];  // first position of a bin in `order`
  int cursor[kColBins];
  unsigned lo_bits, hi_bits, w_bits;  // min / max centre, max width as orderable floats
  float xmin, inv_w, half_wmax;
};

__device__ __forceinline__ int col_bin(const ColBins* cb, const float cx) {
  const float f = (cx - cb->xmin) * cb->inv_w;
  return !(f >= 0.f) ? 0 : (f >= static_cast<float>(kColBins - 1) ? kColBins - 1 : static_cast<int>(f));
}

// All threads call it (block barriers inside).  s_order: m ints.
template <int THREADS>
__device__ __forceinline__ void bin_columns(const float4* s_box, const int* s_list, const int m, ColBins* cb, int* s_order) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    cb->lo_bits = 0xffffffffu;
    cb->hi_bits = 0u;
    cb->w_bits = 0u;
  }
  for (int b = tid; b < kColBins; b += THREADS) cb->cursor[b] = 0;
  __syncthreads();
  for (int j = tid; j < m; j += THREADS) {
    const float4 d = s_box[s_list[j]];
    const float cx = (d.x + d.z) * 0.5f, w = d.z - d.x;
    if (cx == cx && fabsf(cx) < 1e30f) {  // (NaN / infinite boxes have IoU 0 with everything: where they land does not matter)
      atomicMin(&cb->lo_bits, float_orderable(cx));
      atomicMax(&cb->hi_bits, float_orderable(cx));
    }
    if (w == w && w < 1e30f) atomicMax(&cb->w_bits, float_orderable(fmaxf(w, 0.f)));
  }
  __syncthreads();
  if (tid == 0) {
    const float lo = cb->lo_bits == 0xffffffffu ? 0.f : float_from_orderable(cb->lo_bits);
    const float hi = cb->hi_bits == 0u ? 0.f : float_from_orderable(cb->hi_bits);
    cb->xmin = lo;
    cb->inv_w = static_cast<float>(kColBins) / fmaxf(hi - lo, 1e-3f);
    cb->half_wmax = cb->w_bits == 0u ? 0.f : 0.5f * float_from_orderable(cb->w_bits);
  }
  __syncthreads();
  for (int j = tid; j < m; j += THREADS) {
    const float4 d = s_box[s_list[j]];
    atomicAdd(&cb->cursor[col_bin(cb, (d.x + d.z) * 0.5f)], 1);
  }
  __syncthreads();
  if (tid < 32) {  // exclusive prefix over the 64 bins by one warp
    const int c0 = cb->cursor[2 * tid], c1 = cb->cursor[2 * tid + 1];
    int incl = c0 + c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(kFull, incl, d);
      if (tid >= d) incl += o;
    }
    const int base = incl - (c0 + c1);
    cb->start[2 * tid] = base;
    cb->start[2 * tid + 1] = base + c0;
    cb->cursor[2 * tid] = base;
    cb->cursor[2 * tid + 1] = base + c0;
    if (tid == 31) cb->start[kColBins] = incl;
  }
  __syncthreads();
  for (int j = tid; j < m; j += THREADS) {
    const float4 d = s_box[s_list[j]];
    s_order[atomicAdd(&cb->cursor[col_bin(cb, (d.x + d.z) * 0.5f)], 1)] = j;
  }
  __syncthreads();
}

// positions [*p0, *p1) of `order` that can hold a column intersecting box a
__device__ __forceinline__ void col_range(const ColBins* cb, const float4 a, int* p0, int* p1) {
  const int b0 = max(col_bin(cb, a.x - cb->half_wmax) - 1, 0);
  const int b1 = min(col_bin(cb, a.z + cb->half_wmax) + 1, kColBins - 1);
  *p0 = cb->start[b0];
  *p1 = cb->start[b1 + 1];
}

// One association stage: rows = tracks with s_match[t] < 0, columns = s_list[0..m).
// On return s_match[t] holds (det index | flag) for the rows that won their column.
// A group of G lanes (G = 4 .. 32, a power of two chosen from the number of columns) takes a row at
// a time: its lanes split the columns (each keeps the first arg-max of its ascending subsequence),
// then combine - larger IoU wins, equal IoU resolves to the lower column, which is np.argmax over
// the whole row (tracker.py:187).
template <int THREADS>
__device__ __forceinline__ void associate(const TrackPrefetch* pf, const float4* __restrict__ g_box,
                                          const float* kf_mean, const int32_t* g_tsu, size_t row0, int T,
                                          const float4* s_box, const float* s_area,
                                          const int* s_list, int m, int* s_win, int* s_match,
                                          float thresh, int flag, ColBins* cb, int* s_order) {
  const int tid = threadIdx.x;
  for (int j = tid; j < m; j += THREADS) s_win[j] = INT_MAX;
  const bool binned = m >= kBinMinCols && thresh > 0.f;  // block-uniform
  if (binned) bin_columns<THREADS>(s_box, s_list, m, cb, s_order);
  else __syncthreads();
  // lanes per row: few enough that the rows of a typical table (a few hundred) keep all groups busy,
  // enough that a lane's share of the columns stays short
  const int G = m <= 8 ? 4 : (m <= 64 ? 8 : (m <= 256 ? 16 : 32));
  const int sub = tid & (G - 1), groups = THREADS / G;
  // every lane of a warp runs the same number of rounds (shuffles need the whole warp)
  for (int t0 = 0; t0 < T; t0 += groups) {
    const int t = t0 + tid / G;
    const bool open = t < T && s_match[t] < 0;
    float best = -1.f;
    int bj = INT_MAX;
    if (open) {
      const float4 a = t < kTrackPrefRows ? pf->abox[t]
                                          : (kf_mean ? kalman_predicted_box(kf_mean, row0 + t, g_tsu[row0 + t]) : g_box[t]);
      const float area_a = box_area(a);
      if (binned) {
        int p0, p1;
        col_range(cb, a, &p0, &p1);
        for (int p = p0 + sub; p < p1; p += G) {
          const int j = s_order[p], d = s_list[j];
          const float v = pair_iou(a, area_a, s_box[d], s_area[d]);
          if (v > best || (v == best && j < bj)) {  // the bin order is arbitrary: the lower column wins a tie
            best = v;
            bj = j;
          }
        }
      } else {
        for (int j = sub; j < m; j += G) {
          const int d = s_list[j];
          const float v = pair_iou(a, area_a, s_box[d], s_area[d]);
          if (v > best) {  // strict: first arg-max of the lane's columns
            best = v;
            bj = j;
          }
        }
      }
    }
    for (int d = G >> 1; d > 0; d >>= 1) {
      const float ov = __shfl_xor_sync(kFull, best, d);
      const int oj = __shfl_xor_sync(kFull, bj, d);
      if (ov > best || (ov == best && oj < bj)) {
        best = ov;
        bj = oj;
      }
    }
    if (open && sub == 0 && best >= thresh) {  // tracker.py:188, float32 compare
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

// ---------------------------------------------------------------------------------------
// Optimal assignment (RTM_ASSIGN_OPTIMAL): what the reference computes when `lap` is installed,
// lap.lapjv(1 - IoU, extend_cost=True, cost_limit=1 - thresh) (tracker.py:168-181).  lap solves
// the (T + N) x (T + N) problem [[C, L/2], [L/2, 0]] (L = cost_limit): a pair is worth matching
// iff its cost is below L, so the optimum is the minimum-cost matching on the ADMISSIBLE pairs
// (cost < L, i.e. IoU > thresh - a sparse graph) where every unmatched row or column pays L/2.
//   1. admissible pairs are collected in parallel (same row scan as the greedy stage);
//   2. a pair whose row and column have no other admissible pair is matched at once - in ordinary
//      scenes that settles everything;
//   3. what is left falls into small connected components; each is solved exactly with the
//      Hungarian method (potentials + shortest augmenting paths, float64) on its own extended
//      matrix, serially - the components of a stream are few and tiny.
// Limits (reported through the status word, never silently exceeded): kAssignMaxEdges admissible
// pairs per stage, kAssignMaxSide rows or columns per component.
// ---------------------------------------------------------------------------------------
constexpr int kAssignMaxEdges = 4096;
constexpr int kAssignMaxSide = 32;

struct AssignScratch {
  int n_edges;
  int limit_hit;
  int n_open;                                // pairs left after the pairs alone in their row and column are settled
  unsigned short open[kAssignMaxEdges];      // their indices, ascending
  unsigned short edge_row[kAssignMaxEdges], edge_col[kAssignMaxEdges];
  float edge_cost[kAssignMaxEdges];  // 1 - IoU in float32 (what the reference hands to lap), < 0 = settled
  // working set of the component solver (thread 0 only)
  float cost[kAssignMaxSide * kAssignMaxSide];
  int rows[kAssignMaxSide], cols[kAssignMaxSide], match[kAssignMaxSide];
  double u[2 * kAssignMaxSide + 1], v[2 * kAssignMaxSide + 1], minv[2 * kAssignMaxSide + 1];
  int p[2 * kAssignMaxSide + 1], way[2 * kAssignMaxSide + 1];
  unsigned char used[2 * kAssignMaxSide + 8];
};

// Hungarian method on the extended matrix of one component: r rows, c columns, cost[i * c + j]
// (float32 values, +inf where the pair is not admissible), half = L / 2.  match[i] = column or -1.
static __device__ __noinline__ void hungarian_component(AssignScratch* sc, const int r, const int c, const double half) {
  const int n = r + c;
  const double kInf = 1e300;
  const float* cost = sc->cost;
  double *u = sc->u, *v = sc->v, *minv = sc->minv;
  int *p = sc->p, *way = sc->way, *match = sc->match;
  unsigned char* used = sc->used;
  auto entry = [&](int i, int j) -> double {  // 0-based
    if (i < r && j < c) {
      const float x = cost[i * c + j];
      return x < 3.0e38f ? static_cast<double>(x) : kInf;
    }
    if (i >= r && j >= c) return 0.0;
    return half;
  };
  for (int k = 0; k <= n; ++k) {
    u[k] = v[k] = 0.0;
    p[k] = 0;
    way[k] = 0;
  }
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    for (int k = 0; k <= n; ++k) {
      minv[k] = kInf;
      used[k] = 0;
    }
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = kInf;
      int j1 = 0;
      for (int j = 1; j <= n; ++j) {
        if (used[j]) continue;
        const double cur = entry(i0 - 1, j - 1) - u[i0] - v[j];
        if (cur < minv[j]) {
          minv[j] = cur;
          way[j] = j0;
        }
        if (minv[j] < delta) {
          delta = minv[j];
          j1 = j;
        }
      }
      for (int j = 0; j <= n; ++j) {
        if (used[j]) {
          u[p[j]] += delta;
          v[j] -= delta;
        } else {
          minv[j] -= delta;
        }
      }
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0);
  }
  for (int i = 0; i < r; ++i) match[i] = -1;
  for (int j = 1; j <= c; ++j)
    if (p[j] >= 1 && p[j] <= r) match[p[j] - 1] = j - 1;
}

// ---------------------------------------------------------------------------------------
// The general solver: any number of admissible pairs, components of any size, in global scratch.
//
// lap's extended problem [[C, L/2], [L/2, 0]] has the same optima as this one: every row takes an admissible
// column at cost c - L (< 0) or stays unmatched at cost 0; columns are used at most once.  It is solved exactly by
// shortest augmenting paths (Jonker-Volgenant form: row potentials u <= 0, column potentials v, reduced costs
// c - L - u - v >= 0, float64), one row insertion after the other, each a Dijkstra search over the sparse
// adjacency in which the whole block relaxes and selects: "leave this tree row unmatched" is a terminal of its own
// beside "reach a free column".  Ties are resolved towards the lower column / the row reached first, so the result does
// not depend on the order in which the pairs were collected.  Rows are the stage's still-open rows, columns its
// still-free columns: what the shared-memory solver has settled before (pairs alone in their row and column, small
// components) stays settled - components are independent.
// ---------------------------------------------------------------------------------------
struct AssignGlobal {
  int e_cap;
  int* row_start;           // (T + 1)
  int* row_col;             // (T)   column matched to the row, -1 = none
  double* u;                // (T)
  double* v;                // (S)
  double* dist;             // (S)
  int* pred;                // (S)   tree row a column was reached from
  int* col_row;             // (S)   row matched to the column, -1 = free
  int* scanned;             // (S)
  float* e_cost;            // (e_cap) 1 - IoU, float32 as the reference hands it to lap
  unsigned short* e_col;    // (e_cap)
};

inline __host__ __device__ size_t assign_global_fixed_bytes(int capacity, int det_stride) {
  return (static_cast<size_t>(capacity) + 1) * 4 + static_cast<size_t>(capacity) * (4 + 8) + static_cast<size_t>(det_stride) * (8 + 8 + 4 + 4 + 4) + 64;
}

__device__ __forceinline__ AssignGlobal assign_global_layout(unsigned char* base, size_t bytes, int capacity, int S) {
  AssignGlobal g;
  unsigned char* p = base;
  g.u = reinterpret_cast<double*>(p); p += static_cast<size_t>(capacity) * 8;
  g.v = reinterpret_cast<double*>(p); p += static_cast<size_t>(S) * 8;
  g.dist = reinterpret_cast<double*>(p); p += static_cast<size_t>(S) * 8;
  g.row_start = reinterpret_cast<int*>(p); p += (static_cast<size_t>(capacity) + 1) * 4;
  g.row_col = reinterpret_cast<int*>(p); p += static_cast<size_t>(capacity) * 4;
  g.pred = reinterpret_cast<int*>(p); p += static_cast<size_t>(S) * 4;
  g.col_row = reinterpret_cast<int*>(p); p += static_cast<size_t>(S) * 4;
  g.scanned = reinterpret_cast<int*>(p); p += static_cast<size_t>(S) * 4;
  p = base + ((p - base + 15) & ~static_cast<size_t>(15));
  const size_t left = bytes > static_cast<size_t>(p - base) ? bytes - (p - base) : 0;
  g.e_cap = static_cast<int>(left / 6 > 0x3fffffff ? 0x3fffffff : left / 6);
  g.e_cost = reinterpret_cast<float*>(p);
  g.e_col = reinterpret_cast<unsigned short*>(p + static_cast<size_t>(g.e_cap) * 4);
  return g;
}

// Returns 1 when the scratch cannot hold the stage's admissible pairs (nothing is changed then), else 0.
template <int THREADS, typename RowBox>
__device__ __forceinline__ int assign_general(RowBox row_box, const int T, const float4* s_box, const float* s_area,
                                              const int* s_list, const int m, int* s_win, int* s_match, const double cost_limit,
                                              const int flag, const ColBins* cb, const int* s_order, const bool binned, int* deg_row,
                                              int* s_scan, const AssignGlobal g) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = THREADS / 32;
  const double kInf = 1e300;
  __shared__ double r_val[32];
  __shared__ int r_idx[32];
  __shared__ double sh_d;      // distance of the row being scanned
  __shared__ double sh_dummy;  // cheapest "leave a tree row unmatched" so far
  __shared__ int sh_i, sh_dummy_row, sh_state, sh_total;
  const int G = 8, sub = tid & (G - 1), groups = THREADS / G;  // lanes per row while collecting pairs

  auto for_pairs = [&](const int t, auto&& fn) {  // admissible pairs (j, cost) of open row t over the stage's free columns
    const float4 a = row_box(t);
    const float area_a = box_area(a);
    int p0 = 0, p1 = m;
    if (binned) col_range(cb, a, &p0, &p1);
    for (int p = p0 + sub; p < p1; p += G) {
      const int j = binned ? s_order[p] : p, d = s_list[j];
      if (s_win[j] != INT_MAX) continue;
      const float cst = __fsub_rn(1.f, pair_iou(a, area_a, s_box[d], s_area[d]));
      if (static_cast<double>(cst) < cost_limit) fn(j, cst);
    }
  };
  // ---- pairs per open row, CSR offsets ----
  for (int t = tid; t < T; t += THREADS) deg_row[t] = 0;
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += groups) {
    const int t = t0 + tid / G;
    if (t < T && s_match[t] < 0) {
      int n = 0;
      for_pairs(t, [&](int, float) { ++n; });
      if (n) atomicAdd(&deg_row[t], n);
    }
  }
  __syncthreads();
  int run = 0;
  for (int t0 = 0; t0 < T; t0 += THREADS) {
    const int t = t0 + tid;
    int tot;
    const int off = run + block_exclusive_sum<THREADS>(t < T ? deg_row[t] : 0, s_scan, &tot);
    if (t < T) g.row_start[t] = off;
    run += tot;
  }
  if (tid == 0) {
    g.row_start[T] = run;
    sh_total = run;
  }
  __syncthreads();
  const int E = sh_total;
  if (E > g.e_cap) return 1;
  for (int t = tid; t < T; t += THREADS) deg_row[t] = 0;  // now the fill cursor of the row
  for (int j = tid; j < m; j += THREADS) {
    g.v[j] = 0.0;
    g.col_row[j] = -1;
  }
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += groups) {
    const int t = t0 + tid / G;
    if (t < T && s_match[t] < 0) {
      const int base = g.row_start[t];
      for_pairs(t, [&](int j, float cst) {
        const int k = base + atomicAdd(&deg_row[t], 1);
        g.e_col[k] = static_cast<unsigned short>(j);
        g.e_cost[k] = cst;
      });
    }
  }
  __syncthreads();
  // u = the row's cheapest shifted cost (<= 0): every reduced cost starts non-negative
  for (int t = tid; t < T; t += THREADS) {
    double lo = 0.0;
    for (int k = g.row_start[t]; k < g.row_start[t + 1]; ++k) lo = fmin(lo, static_cast<double>(g.e_cost[k]) - cost_limit);
    g.u[t] = lo;
    g.row_col[t] = -1;
  }
  __syncthreads();
  // ---- one augmentation per open row that has a pair ----
  for (int i0 = 0; i0 < T; ++i0) {
    if (g.row_start[i0] == g.row_start[i0 + 1]) continue;  // block-uniform (global memory written before the barrier above)
    for (int j = tid; j < m; j += THREADS) {
      g.dist[j] = kInf;
      g.scanned[j] = 0;
    }
    if (tid == 0) {
      sh_i = i0;
      sh_d = 0.0;
      sh_dummy = kInf;
      sh_dummy_row = -1;
      sh_state = 0;
    }
    __syncthreads();
    while (true) {
      const int i = sh_i;
      const double di = sh_d, ui = g.u[i];
      // relax the row's pairs (few: one warp is plenty); note the cheaper way out through the row itself
      if (warp == 0) {
        for (int k = g.row_start[i] + lane; k < g.row_start[i + 1]; k += 32) {
          const int j = g.e_col[k];
          if (g.scanned[j]) continue;
          const double nd = di + ((static_cast<double>(g.e_cost[k]) - cost_limit) - ui - g.v[j]);
          if (nd < g.dist[j]) {
            g.dist[j] = nd;
            g.pred[j] = i;
          }
        }
        if (lane == 0) {
          const double dd = di - ui;  // the row's own way out: reduced cost 0 - u_i
          if (dd < sh_dummy) {
            sh_dummy = dd;
            sh_dummy_row = i;
          }
        }
      }
      __syncthreads();
      // nearest unscanned column (ties: the lower column)
      double best = kInf;
      int bj = INT_MAX;
      for (int j = tid; j < m; j += THREADS) {
        const double dj = g.dist[j];
        if (!g.scanned[j] && (dj < best || (dj == best && j < bj))) {
          best = dj;
          bj = j;
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(kFull, best, d);
        const int oj = __shfl_xor_sync(kFull, bj, d);
        if (ob < best || (ob == best && oj < bj)) {
          best = ob;
          bj = oj;
        }
      }
      if (lane == 0) {
        r_val[warp] = best;
        r_idx[warp] = bj;
      }
      __syncthreads();
      if (warp == 0) {
        best = lane < kWarps ? r_val[lane] : kInf;
        bj = lane < kWarps ? r_idx[lane] : INT_MAX;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const double ob = __shfl_xor_sync(kFull, best, d);
          const int oj = __shfl_xor_sync(kFull, bj, d);
          if (ob < best || (ob == best && oj < bj)) {
            best = ob;
            bj = oj;
          }
        }
        if (lane == 0) {
          if (!(best < sh_dummy)) {
            sh_state = 1;  // the path ends in a tree row that stays unmatched
          } else if (g.col_row[bj] < 0) {
            sh_state = 2;  // ... or in a free column
            r_idx[0] = bj;
            sh_d = best;
          } else {
            g.scanned[bj] = 1;
            sh_i = g.col_row[bj];
            sh_d = best;
          }
        }
      }
      __syncthreads();
      if (sh_state) break;
    }
    // potentials: every tree node moves by (final distance - its own distance)
    const double D = sh_state == 1 ? sh_dummy : sh_d;
    for (int j = tid; j < m; j += THREADS) {
      if (g.scanned[j]) {
        const double delta = D - g.dist[j];
        g.v[j] -= delta;
        g.u[g.col_row[j]] += delta;
      }
    }
    if (tid == 0) g.u[i0] += D;
    __syncthreads();
    if (tid == 0) {  // flip the path
      int j;
      if (sh_state == 1) {
        const int ik = sh_dummy_row;
        j = g.row_col[ik];
        g.row_col[ik] = -1;
        if (ik == i0) j = -1;
      } else {
        j = r_idx[0];
      }
      while (j >= 0) {
        const int i = g.pred[j];
        const int jn = g.row_col[i];
        g.row_col[i] = j;
        g.col_row[j] = i;
        j = i == i0 ? -1 : jn;
      }
    }
    __syncthreads();
  }
  for (int t = tid; t < T; t += THREADS) {
    const int j = g.row_col[t];
    if (s_match[t] < 0 && g.row_start[t] != g.row_start[t + 1] && j >= 0) {
      s_match[t] = s_list[j] | flag;
      s_win[j] = t;
    }
  }
  __syncthreads();
  return 0;
}

template <int THREADS>
__device__ __forceinline__ int associate_optimal(const TrackPrefetch* pf, const float4* __restrict__ g_box,
                                                 const float* kf_mean, const int32_t* g_tsu, size_t row0, int T,
                                                 const float4* s_box, const float* s_area, const int* s_list, int m,
                                                 int* s_win, int* s_match, const double cost_limit, int flag,
                                                 AssignScratch* sc, int* deg_row, int* deg_col, ColBins* cb, int* s_order,
                                                 int* s_scan, unsigned char* scratch, size_t scratch_bytes, int capacity, int det_stride) {
  const int tid = threadIdx.x;
  // admissible pairs have IoU > match_thresh: with a positive threshold (cost_limit < 1) they intersect
  const bool binned = m >= kBinMinCols && cost_limit < 1.0;  // block-uniform
  if (binned) bin_columns<THREADS>(s_box, s_list, m, cb, s_order);
  for (int j = tid; j < m; j += THREADS) {
    s_win[j] = INT_MAX;
    deg_col[j] = 0;
  }
  for (int t = tid; t < T; t += THREADS) deg_row[t] = 0;
  if (tid == 0) {
    sc->n_edges = 0;
    sc->limit_hit = 0;
  }
  __syncthreads();
  // ---- 1. admissible pairs ----
  const int G = m <= 8 ? 4 : (m <= 64 ? 8 : (m <= 256 ? 16 : 32));
  const int sub = tid & (G - 1), groups = THREADS / G;
  for (int t0 = 0; t0 < T; t0 += groups) {
    const int t = t0 + tid / G;
    if (t < T && s_match[t] < 0) {
      const float4 a = t < kTrackPrefRows ? pf->abox[t]
                                          : (kf_mean ? kalman_predicted_box(kf_mean, row0 + t, g_tsu[row0 + t]) : g_box[t]);
      const float area_a = box_area(a);
      int p0 = 0, p1 = m;
      if (binned) col_range(cb, a, &p0, &p1);
      for (int p = p0 + sub; p < p1; p += G) {
        const int j = binned ? s_order[p] : p, d = s_list[j];
        const float cst = __fsub_rn(1.f, pair_iou(a, area_a, s_box[d], s_area[d]));  // tracker.py:167, float32
        if (static_cast<double>(cst) < cost_limit) {
          const int e = atomicAdd(&sc->n_edges, 1);
          if (e < kAssignMaxEdges) {
            sc->edge_row[e] = static_cast<unsigned short>(t);
            sc->edge_col[e] = static_cast<unsigned short>(j);
            sc->edge_cost[e] = cst;
          }
          atomicAdd(&deg_row[t], 1);
          atomicAdd(&deg_col[j], 1);
        }
      }
    }
  }
  __syncthreads();
  const int E = min(sc->n_edges, kAssignMaxEdges);
  // ---- 2. pairs alone in their row and column ----
  for (int e = tid; e < E; e += THREADS) {
    const int t = sc->edge_row[e], j = sc->edge_col[e];
    if (deg_row[t] == 1 && deg_col[j] == 1) {
      s_match[t] = s_list[j] | flag;
      s_win[j] = t;
      sc->edge_cost[e] = -1.f;
    }
  }
  __syncthreads();
  // ---- 3. the rest, component by component ----
  // The open pairs are listed in index order (in parallel); thread 0 then floods one component after the other from the
  // first open pair, in the order a scan over all pairs would: membership of a row / column is a stamp in deg_row /
  // deg_col (component number and local index, degrees are not needed any more), and a finished component's pairs
  // leave the list - the walk over "every unsettled pair from e0 on" shrinks with every component instead of
  // re-reading the settled ones and comparing against the member lists.
  {
    int run = 0;
    for (int e0 = 0; e0 < E; e0 += THREADS) {
      const int e = e0 + tid;
      const bool is_open = e < E && !(sc->edge_cost[e] < 0.f);
      int tot;
      const int p = run + block_exclusive_count<THREADS>(is_open, s_scan, &tot);
      if (is_open) sc->open[p] = static_cast<unsigned short>(e);
      run += tot;
    }
    if (tid == 0) sc->n_open = run;
    __syncthreads();
  }
  if (tid == 0) {
    if (sc->n_edges > kAssignMaxEdges) sc->limit_hit = 1;  // the list is incomplete: components cannot be told from it
    int U = sc->n_open;
    for (int comp = 1; U > 0 && !sc->limit_hit; ++comp) {
      int *rows = sc->rows, *cols = sc->cols, r = 0, c = 0;
      float* cost = sc->cost;
      bool too_big = false;
      const int base = -(comp * 64) - 1;  // member k of this component carries base - k (kAssignMaxSide <= 64)
      auto local = [&](const int mark) { return (mark <= base && mark > base - 64) ? base - mark : -1; };
      const int e_first = sc->open[0];
      rows[r] = sc->edge_row[e_first];
      deg_row[rows[r]] = base - r;
      ++r;
      cols[c] = sc->edge_col[e_first];
      deg_col[cols[c]] = base - c;
      ++c;
      // flood: pull in every open pair that shares a row or a column with the component
      for (bool grown = true; grown && !too_big;) {
        grown = false;
        for (int k = 0; k < U && !too_big; ++k) {
          const int e = sc->open[k];
          const int t = sc->edge_row[e], j = sc->edge_col[e];
          const bool has_r = local(deg_row[t]) >= 0, has_c = local(deg_col[j]) >= 0;
          if (has_r == has_c) continue;  // neither (not ours) or both (already in)
          if (!has_r) {
            if (r == kAssignMaxSide) too_big = true;
            else {
              rows[r] = t;
              deg_row[t] = base - r;
              ++r;
            }
          } else {
            if (c == kAssignMaxSide) too_big = true;
            else {
              cols[c] = j;
              deg_col[j] = base - c;
              ++c;
            }
          }
          grown = true;
        }
      }
      if (too_big) {
        sc->limit_hit = 1;
        break;
      }
      for (int k = 0; k < r * c; ++k) cost[k] = 3.4e38f;
      int kept = 0;
      for (int k = 0; k < U; ++k) {
        const int e = sc->open[k];
        const int li = local(deg_row[sc->edge_row[e]]), lj = local(deg_col[sc->edge_col[e]]);
        if (li >= 0 && lj >= 0) {
          cost[li * c + lj] = sc->edge_cost[e];
          sc->edge_cost[e] = -1.f;  // settled with this component
        } else {
          sc->open[kept++] = static_cast<unsigned short>(e);
        }
      }
      U = kept;
      hungarian_component(sc, r, c, cost_limit * 0.5);
      const int* match = sc->match;
      for (int k = 0; k < r; ++k) {
        if (match[k] >= 0 && cost[k * c + match[k]] < 3.0e38f) {
          s_match[rows[k]] = s_list[cols[match[k]]] | flag;
          s_win[cols[match[k]]] = rows[k];
        }
      }
    }
  }
  __syncthreads();
  if (!sc->limit_hit) return 0;
  // beyond the shared-memory solver: everything still open goes to the general one, if the caller gave it scratch
  if (!scratch || scratch_bytes < assign_global_fixed_bytes(capacity, det_stride)) return 1;
  auto row_box = [&](const int t) {
    return t < kTrackPrefRows ? pf->abox[t] : (kf_mean ? kalman_predicted_box(kf_mean, row0 + t, g_tsu[row0 + t]) : g_box[t]);
  };
  return assign_general<THREADS>(row_box, T, s_box, s_area, s_list, m, s_win, s_match, cost_limit, flag, cb, s_order, binned, deg_row,
                                 s_scan, assign_global_layout(scratch, scratch_bytes, capacity, det_stride));
}

// One stream.  `smem_raw`: track_smem_bytes(det_stride, capacity) bytes of shared memory,
// 16-byte aligned.  All THREADS threads of the block must call it (block-uniform control flow).
template <int THREADS, bool WITH_OPTIMAL = false>
__device__ __forceinline__ void track_stream(const TrackArgs& a, const int b, unsigned char* smem_raw,
                                             const TrackPrefetch* pf) {
  const int tid = threadIdx.x;
  const int cap = a.tin.capacity, S = a.det_stride;

  float4* s_box = reinterpret_cast<float4*>(smem_raw);  // S
  float* s_area = reinterpret_cast<float*>(s_box + S);  // S
  int* s_hi = reinterpret_cast<int*>(s_area + S);       // S  det index of the j-th high det
  int* s_lo = s_hi + S;                                 // S  ... low det; later: birth list
  int* s_win = s_lo + S;                                // S  column winners of a stage
  int* s_born = s_win + S;                              // S  1 if high column j is unmatched
  float* s_conf = reinterpret_cast<float*>(s_born + S);  // S
  int* s_cls = reinterpret_cast<int*>(s_conf + S);       // S
  int* s_match = s_cls + S;                             // cap
  int* s_scan = s_match + cap;                          // 33 (+ 7 spare)
  int* s_order = s_scan + 40;                           // S  columns of a stage in bin order
  ColBins* s_bins = reinterpret_cast<ColBins*>(s_order + S);
  // optimal-assignment scratch (only laid out by track_smem_bytes when that mode is on)
  int* s_deg_row = reinterpret_cast<int*>(s_bins + 1);  // cap
  int* s_deg_col = s_deg_row + cap;                     // S
  AssignScratch* s_assign = reinterpret_cast<AssignScratch*>((reinterpret_cast<uintptr_t>(s_deg_col + S) + 15) & ~static_cast<uintptr_t>(15));
  const bool optimal = WITH_OPTIMAL && a.assignment == RTM_ASSIGN_OPTIMAL;
  unsigned char* scratch = a.assign_scratch ? a.assign_scratch + static_cast<size_t>(b) * a.assign_scratch_per_stream : nullptr;

  const size_t row0 = static_cast<size_t>(b) * cap;
  const size_t det0 = static_cast<size_t>(b) * S;
  const float4* in_box = reinterpret_cast<const float4*>(a.tin.xyxy) + row0;
  float4* out_box = reinterpret_cast<float4*>(a.tout.xyxy) + row0;
  const float4* det_box = reinterpret_cast<const float4*>(a.det_xyxy) + det0;

  const int T = pf->count;
  const int next_id = pf->next_id;
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
      const bool p = t < kTrackPrefRows;
      a.tout.track_id[row0 + t] = p ? pf->track_id[t] : a.tin.track_id[row0 + t];
      out_box[t] = p ? pf->box[t] : in_box[t];
      a.tout.confidence[row0 + t] = p ? pf->confidence[t] : a.tin.confidence[row0 + t];
      a.tout.class_id[row0 + t] = p ? pf->class_id[t] : a.tin.class_id[row0 + t];
      a.tout.age[row0 + t] = p ? pf->age[t] : a.tin.age[row0 + t];
      const int tsu_in = p ? pf->tsu[t] : a.tin.time_since_update[row0 + t];
      a.tout.time_since_update[row0 + t] = tsu_in + 1;
      if (a.src_row) a.src_row[row0 + t] = t;
      if (a.kf_mean_in) {  // the filter still advances one frame
        KalmanTrack k;
        kalman_load(k, a.kf_mean_in, a.kf_cov_in, row0 + t);
        kalman_predict(k, tsu_in);
        kalman_store(k, a.kf_mean_out, a.kf_cov_out, row0 + t);
      }
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
      const float cf = a.det_conf[det0 + d];
      s_conf[d] = cf;
      s_cls[d] = a.det_cls[det0 + d];
      hi = cf >= a.track_thresh;
      if (a.det_track_id) a.det_track_id[det0 + d] = 0;
      if (a.det_kind) a.det_kind[det0 + d] = RTM_DET_NONE;
    }
    int tot;
    const int p = block_exclusive_count<THREADS>(hi, s_scan, &tot);
    if (hi) s_hi[H + p] = d;
    if (valid && !hi) s_lo[L + (tid - p)] = d;
    const int in_round = min(THREADS, n - r0);
    H += tot;
    L += in_round - tot;
  }
  for (int t = tid; t < T; t += THREADS) s_match[t] = -1;
  for (int j = tid; j < H; j += THREADS) s_born[j] = 1;
  __syncthreads();
  RTM_TL(11);

  // ---- stage 1: all retained tracks x high detections (tracker.py:91-104) ---------------
  if (T > 0 && H > 0) {
    if (WITH_OPTIMAL && optimal) {
      if (associate_optimal<THREADS>(pf, in_box, a.kf_mean_in, a.tin.time_since_update, row0, T, s_box, s_area, s_hi, H, s_win,
                                     s_match, a.cost_limit, 0, s_assign, s_deg_row, s_deg_col, s_bins, s_order, s_scan, scratch, a.assign_scratch_per_stream, cap, S))
        st |= RTM_STATUS_ASSIGN_LIMIT;
    } else {
      associate<THREADS>(pf, in_box, a.kf_mean_in, a.tin.time_since_update, row0, T, s_box, s_area, s_hi, H, s_win, s_match,
                         a.match_thresh, 0, s_bins, s_order);
    }
    for (int j = tid; j < H; j += THREADS) s_born[j] = (s_win[j] == INT_MAX);
    __syncthreads();
  }
  RTM_TL(12);
  // ---- stage 2: still-unmatched tracks x low detections, same threshold (tracker.py:109-123)
  if (T > 0 && L > 0) {
    if (WITH_OPTIMAL && optimal) {
      if (associate_optimal<THREADS>(pf, in_box, a.kf_mean_in, a.tin.time_since_update, row0, T, s_box, s_area, s_lo, L, s_win,
                                     s_match, a.cost_limit, kStage2Flag, s_assign, s_deg_row, s_deg_col, s_bins, s_order, s_scan, scratch, a.assign_scratch_per_stream, cap, S))
        st |= RTM_STATUS_ASSIGN_LIMIT;
    } else {
      associate<THREADS>(pf, in_box, a.kf_mean_in, a.tin.time_since_update, row0, T, s_box, s_area, s_lo, L, s_win, s_match,
                         a.match_thresh, kStage2Flag, s_bins, s_order);
    }
  }

  RTM_TL(13);
  // ---- births: unmatched high detections in ascending order (tracker.py:126-135) --------
  int NB = 0;
  int* s_birth = s_lo;  // low list is dead from here on
  for (int r0 = 0; r0 < H; r0 += THREADS) {
    const int j = r0 + tid;
    const bool born = j < H && s_born[j];
    int tot;
    const int p = block_exclusive_count<THREADS>(born, s_scan, &tot);
    if (born) s_birth[NB + p] = s_hi[j];
    NB += tot;
  }
  __syncthreads();

  RTM_TL(14);
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
    KalmanTrack kf;
    if (v < T) {
      const bool p = v < kTrackPrefRows;
      if (a.kf_mean_in) {
        kalman_load(kf, a.kf_mean_in, a.kf_cov_in, row0 + v);
        kalman_predict(kf, p ? pf->tsu[v] : a.tin.time_since_update[row0 + v]);
      }
      id = p ? pf->track_id[v] : a.tin.track_id[row0 + v];
      const int m = s_match[v];
      if (m >= 0) {
        det = m & ~kStage2Flag;
        kind = (m & kStage2Flag) ? RTM_DET_STAGE2 : RTM_DET_STAGE1;
        box = s_box[det];
        conf = s_conf[det];
        cls = s_cls[det];
        age = (p ? pf->age[v] : a.tin.age[row0 + v]) + 1;
        tsu = 1;
        if (a.kf_mean_in) kalman_update(kf, box);
      } else {
        box = p ? pf->box[v] : in_box[v];
        conf = p ? pf->confidence[v] : a.tin.confidence[row0 + v];
        cls = p ? pf->class_id[v] : a.tin.class_id[row0 + v];
        age = p ? pf->age[v] : a.tin.age[row0 + v];
        tsu = (p ? pf->tsu[v] : a.tin.time_since_update[row0 + v]) + 1;
      }
      keep = tsu <= a.track_buffer;
    } else if (v < V) {
      det = s_birth[v - T];
      kind = RTM_DET_BIRTH;
      id = next_id + (v - T);
      box = s_box[det];
      conf = s_conf[det];
      cls = s_cls[det];
      age = 1;
      tsu = 1;
      keep = birth_survives;
      if (a.kf_mean_in) kalman_initiate(kf, box);
    }
    if (det >= 0) {
      if (a.det_track_id) a.det_track_id[det0 + det] = id;
      if (a.det_kind) a.det_kind[det0 + det] = kind;
    }
    int tot;
    const int p = kept + block_exclusive_count<THREADS>(keep, s_scan, &tot);
    if (keep && p < cap) {
      a.tout.track_id[row0 + p] = id;
      out_box[p] = box;
      a.tout.confidence[row0 + p] = conf;
      a.tout.class_id[row0 + p] = cls;
      a.tout.age[row0 + p] = age;
      a.tout.time_since_update[row0 + p] = tsu;
      if (a.src_row) a.src_row[row0 + p] = v < T ? v : -1;
      if (a.kf_mean_in) kalman_store(kf, a.kf_mean_out, a.kf_cov_out, row0 + p);
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

inline size_t track_smem_bytes(int det_stride, int capacity, bool optimal = false) {
  if (optimal)
    return static_cast<size_t>(det_stride) * (16 + 4 + 4 * 4 + 4 + 4 + 4 + 4 + 4) + static_cast<size_t>(capacity) * 8 + 40 * 4 +
           sizeof(ColBins) + sizeof(AssignScratch) + 16;
  return static_cast<size_t>(det_stride) * (16 + 4 + 4 * 4 + 4 + 4 + 4 + 4) + static_cast<size_t>(capacity) * 4 + 40 * 4 + sizeof(ColBins);
}

}  // namespace rtm
