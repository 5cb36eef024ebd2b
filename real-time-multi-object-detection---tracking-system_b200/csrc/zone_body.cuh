// Z1..Z3 for one stream as a device function (used by zone_step_kernel and by the fused
// post-backbone kernel).  See zone.cu for the semantics and the reference citations.
#pragma once

#include <limits.h>
#include <math.h>

#include "rtm_common.cuh"

namespace rtm {

constexpr int kMaxZonesPerStream = 64;  // fired-zone bitmask is one 64-bit word per track

struct ZoneArgs {
  rtm_zone_set zs;
  rtm_track_table trk;
  const int32_t* src_row;
  rtm_zone_state sin, sout;
  double now;
  const double* now_per_stream;
  int32_t frame_id;
  rtm_zone_event* events;
  int32_t event_stride;
  int32_t* event_count;
  int32_t* status;
};

// OpenCV pointPolygonTest, measureDist = false, integer contour and point.
// Returns +1 inside, 0 on the boundary, -1 outside.
__device__ __forceinline__ int point_in_polygon(const int2* __restrict__ poly, int k, int px, int py) {
  if (k <= 0) return -1;
  int counter = 0;
  int2 v = poly[k - 1];
  for (int i = 0; i < k; ++i) {
    const int2 v0 = v;
    v = poly[i];
    if ((v0.y <= py && v.y <= py) || (v0.y > py && v.y > py) || (v0.x < px && v.x < px)) {
      if (py == v.y && (px == v.x || (py == v0.y && ((v0.x <= px && px <= v.x) || (v.x <= px && px <= v0.x)))))
        return 0;
      continue;
    }
    long long d = static_cast<long long>(py - v0.y) * (v.x - v0.x) -
                  static_cast<long long>(px - v0.x) * (v.y - v0.y);
    if (d == 0) return 0;
    if (v.y < v0.y) d = -d;
    counter += d > 0;
  }
  return (counter & 1) ? 1 : -1;
}

// The stream's zone table and polygons, staged in shared memory ahead of use (the fused kernel
// issues these loads before the NMS stage).  kZonePrefVertices vertices are staged; streams with
// more read their polygons from global memory.
#ifndef RTM_ZONE_PREF_VERTICES
#define RTM_ZONE_PREF_VERTICES 512
#endif
constexpr int kZonePrefVertices = RTM_ZONE_PREF_VERTICES;
struct ZonePrefetch {
  int2 poly[kZonePrefVertices];
  double dwell[kMaxZonesPerStream], cool[kMaxZonesPerStream];
  int off[kMaxZonesPerStream + 1], col[kMaxZonesPerStream];
  int4 bbox[kMaxZonesPerStream];  // x_min, y_min, x_max, y_max of the zone's vertices: a point outside is outside the polygon
  int nz, v0, staged;
};

// All threads call it; contains one block barrier; the caller provides another before zone_stream.
template <int THREADS>
__device__ __forceinline__ void zone_prefetch(const ZoneArgs& a, const int b, ZonePrefetch* zp) {
  const int tid = threadIdx.x;
  const int z0 = a.zs.zone_offsets[b], z1 = a.zs.zone_offsets[b + 1];
  const int nz = min(z1 - z0, kMaxZonesPerStream);
  if (tid == 0 && z1 - z0 > kMaxZonesPerStream && a.status) atomicOr(&a.status[b], RTM_STATUS_ZONE_LIMIT);
  const int v0 = a.zs.poly_offsets[z0];
  for (int z = tid; z <= nz; z += THREADS) zp->off[z] = a.zs.poly_offsets[z0 + z] - v0;
  for (int z = tid; z < nz; z += THREADS) {
    zp->col[z] = a.zs.column[z0 + z];
    zp->dwell[z] = a.zs.dwell_sec[z0 + z];
    zp->cool[z] = a.zs.cooldown_sec[z0 + z];
  }
  __syncthreads();
  const int nv = zp->off[nz];
  const bool staged = nv <= kZonePrefVertices;
  const int2* g_poly = reinterpret_cast<const int2*>(a.zs.poly_xy);
  if (staged)
    for (int i = tid; i < nv; i += THREADS) zp->poly[i] = g_poly[v0 + i];
  for (int z = tid; z < nz; z += THREADS) {  // (from global memory: the staged copy is not visible to this thread yet)
    int4 bb = make_int4(INT_MAX, INT_MAX, INT_MIN, INT_MIN);
    for (int i = zp->off[z]; i < zp->off[z + 1]; ++i) {
      const int2 v = g_poly[v0 + i];
      bb.x = min(bb.x, v.x);
      bb.y = min(bb.y, v.y);
      bb.z = max(bb.z, v.x);
      bb.w = max(bb.w, v.y);
    }
    zp->bbox[z] = bb;
  }
  if (tid == 0) {
    zp->nz = nz;
    zp->v0 = v0;
    zp->staged = staged;
  }
}

constexpr int kZoneStepEvents = 4096;  // events one stream can emit in one step (shared-memory staging)

struct ZoneFired {  // a fired (row, zone) pair waiting to be ranked
  int32_t row, zone;
  double dwell;
};

// One stream.  `smem_raw`: zone_smem_bytes(event_stride) bytes of shared memory, `s_scan`: 33 ints.
// All THREADS threads of the block must call it.
//
// Work is spread over (track row, state column) pairs: zones of a stream that share a name share
// a column and are evaluated in order by the pair's thread (their order matters: the reference
// keys its dicts by name); different columns never interact.  Fired (row, zone) pairs are staged
// in shared memory and ranked, so events come out in the reference's (track order, zone order).
template <int THREADS>
__device__ __forceinline__ void zone_stream(const ZoneArgs& a, const int b, unsigned char* smem_raw, int* s_scan,
                                            const ZonePrefetch* zp) {
  const int tid = threadIdx.x;
  const int cap = a.trk.capacity, C = a.zs.num_columns;
  const int nz = zp->nz;
  const int2* poly = zp->staged ? zp->poly : reinterpret_cast<const int2*>(a.zs.poly_xy) + zp->v0;
  ZoneFired* s_ev = reinterpret_cast<ZoneFired*>(smem_raw);
  const int ev_cap = min(a.event_stride, kZoneStepEvents);
  int* s_nev = s_scan;  // one counter
  if (tid == 0) *s_nev = 0;
  __syncthreads();
  RTM_TL(21);

  const double now = a.now_per_stream ? a.now_per_stream[b] : a.now;
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  const size_t row0 = static_cast<size_t>(b) * cap;
  const size_t st0 = static_cast<size_t>(b) * C * cap;  // state is (B, C, capacity)
  const int T = min(a.trk.count[b], cap);
  const float4* boxes = reinterpret_cast<const float4*>(a.trk.xyxy) + row0;
  rtm_zone_event* ev_out = a.events + static_cast<size_t>(b) * a.event_stride;

  // ---- pass 1: (column, row) pairs, column-major so that state accesses coalesce over rows ----
  // (the lanes of a warp are brought back together after every pair: their polygon walks take different turns, and
  // without the barrier the warp drifts apart for the rest of the loop - ncu showed 4 of 32 lanes active, 7x the time)
  const int pairs = T * C;
  for (int q0 = 0; q0 < pairs; q0 += THREADS, __syncwarp()) {
    const int q = q0 + tid;
    if (q >= pairs) continue;
    const int c = q / T, r = q - c * T;
    const int src = a.src_row ? a.src_row[row0 + r] : r;
    const bool active = a.trk.time_since_update[row0 + r] == 1;
    // rows absent from this call lose their dwell timers and keep their cooldowns (zone_engine.py:128-130)
    double fs = kNaN, la = 0.0;
    if (src >= 0) {
      la = a.sin.last_alert[st0 + static_cast<size_t>(c) * cap + src];
      if (active) fs = a.sin.first_seen[st0 + static_cast<size_t>(c) * cap + src];
    }
    if (active) {
      const float4 box = boxes[r];
      const int cx = __float2int_rz(__fdiv_rn(__fadd_rn(box.x, box.z), 2.0f));
      const int cy = __float2int_rz(__fdiv_rn(__fadd_rn(box.y, box.w), 2.0f));
      for (int z = 0; z < nz; ++z) {
        if (zp->col[z] != c) continue;
        const int p0 = zp->off[z], k = zp->off[z + 1] - p0;
        const int4 bb = zp->bbox[z];
        // cv2.pointPolygonTest >= 0 (inside or on the boundary) is impossible outside the vertices' bounding box
        if (cx >= bb.x && cx <= bb.z && cy >= bb.y && cy <= bb.w && point_in_polygon(poly + p0, k, cx, cy) >= 0) {
          if (fs != fs) fs = now;  // not in the zone before: start the dwell timer
          const double dwell = now - fs;
          if (dwell >= zp->dwell[z] && now - la >= zp->cool[z]) {
            la = now;
            const int slot = atomicAdd(s_nev, 1);
            if (slot < ev_cap) s_ev[slot] = ZoneFired{r, z, dwell};
          }
        } else {
          fs = kNaN;
        }
      }
    }
    a.sout.first_seen[st0 + static_cast<size_t>(c) * cap + r] = fs;
    a.sout.last_alert[st0 + static_cast<size_t>(c) * cap + r] = la;
  }
  __syncthreads();

  // ---- pass 2: rank the fired pairs by (row, zone) and write the event records ----
  const int fired_total = *s_nev;
  const int E = min(fired_total, ev_cap);
  for (int e = tid; e < E; e += THREADS) {
    const ZoneFired f = s_ev[e];
    const int key = f.row * kMaxZonesPerStream + f.zone;
    int pos = 0;
    for (int j = 0; j < E; ++j) pos += (s_ev[j].row * kMaxZonesPerStream + s_ev[j].zone) < key;
    const int r = f.row;
    const float4 box = boxes[r];
    rtm_zone_event ev;
    ev.stream = b;
    ev.frame_id = a.frame_id;
    ev.track_id = a.trk.track_id[row0 + r];
    ev.zone = f.zone;
    ev.class_id = a.trk.class_id[row0 + r];
    ev.cx = __float2int_rz(__fdiv_rn(__fadd_rn(box.x, box.z), 2.0f));
    ev.cy = __float2int_rz(__fdiv_rn(__fadd_rn(box.y, box.w), 2.0f));
    ev.row = r;
    ev.dwell = f.dwell;
    ev.now = now;
    ev.xyxy[0] = box.x;
    ev.xyxy[1] = box.y;
    ev.xyxy[2] = box.z;
    ev.xyxy[3] = box.w;
    ev_out[pos] = ev;
  }
  if (tid == 0) {
    a.event_count[b] = E;
    if (fired_total > ev_cap && a.status) atomicOr(&a.status[b], RTM_STATUS_EVENT_OVERFLOW);
  }
}

inline size_t zone_smem_bytes(int event_stride) {
  return static_cast<size_t>(event_stride < kZoneStepEvents ? event_stride : kZoneStepEvents) * sizeof(ZoneFired);
}

}  // namespace rtm
