// Z1..Z3 for one stream as a device function (used by zone_step_kernel and by the fused
// post-backbone kernel).  See zone.cu for the semantics and the reference citations.
#pragma once

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
  int32_t max_vertices;  // shared-memory polygon tile (vertices); larger streams read global
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

// One stream.  `smem_raw`: a.max_vertices * 8 bytes of shared memory; `s_scan`: 33 ints.
// All THREADS threads of the block must call it.
template <int THREADS>
__device__ __forceinline__ void zone_stream(const ZoneArgs& a, const int b, unsigned char* smem_raw, int* s_scan) {
  const int tid = threadIdx.x;
  constexpr int kZoneThreads = THREADS;
  // per-stream zone table staged once (one global round trip instead of several per zone test)
  __shared__ int s_zoff[kMaxZonesPerStream + 1];
  __shared__ int s_zcol[kMaxZonesPerStream];
  __shared__ double s_zdwell[kMaxZonesPerStream];
  __shared__ double s_zcool[kMaxZonesPerStream];
  const int cap = a.trk.capacity, C = a.zs.num_columns;
  const int z0 = a.zs.zone_offsets[b], z1 = a.zs.zone_offsets[b + 1];
  const int nz = min(z1 - z0, kMaxZonesPerStream);
  if (tid == 0 && z1 - z0 > kMaxZonesPerStream && a.status) atomicOr(&a.status[b], RTM_STATUS_ZONE_LIMIT);
  const int v0 = a.zs.poly_offsets[z0];
  if (tid <= nz) s_zoff[tid] = a.zs.poly_offsets[z0 + tid] - v0;
  if (tid < nz) {
    s_zcol[tid] = a.zs.column[z0 + tid];
    s_zdwell[tid] = a.zs.dwell_sec[z0 + tid];
    s_zcool[tid] = a.zs.cooldown_sec[z0 + tid];
  }
  __syncthreads();
  const int nv = s_zoff[nz];

  // stage the stream's polygons (falls back to global memory when they do not fit)
  int2* s_poly = reinterpret_cast<int2*>(smem_raw);
  const int2* g_poly = reinterpret_cast<const int2*>(a.zs.poly_xy);
  const bool staged = nv <= a.max_vertices;
  if (staged)
    for (int i = tid; i < nv; i += kZoneThreads) s_poly[i] = g_poly[v0 + i];
  __syncthreads();
  const int2* poly = staged ? s_poly : g_poly + v0;

  const double now = a.now_per_stream ? a.now_per_stream[b] : a.now;
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  const size_t row0 = static_cast<size_t>(b) * cap;
  const size_t st0 = static_cast<size_t>(b) * C * cap;  // state is (B, C, capacity)
  const int T = min(a.trk.count[b], cap);
  const float4* boxes = reinterpret_cast<const float4*>(a.trk.xyxy) + row0;
  rtm_zone_event* ev_out = a.events + static_cast<size_t>(b) * a.event_stride;

  int ev_base = 0;
  bool overflow = false;
  for (int r0 = 0; r0 < T; r0 += kZoneThreads) {
    const int r = r0 + tid;
    const bool live = r < T;
    unsigned long long fired = 0ull;
    double dwell_of[kMaxZonesPerStream];
    int cx = 0, cy = 0;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const int src = a.src_row ? a.src_row[row0 + r] : r;
      const bool active = a.trk.time_since_update[row0 + r] == 1;
      if (active) box = boxes[r];
      // the row's state, all columns loaded before anything is stored (independent loads in
      // flight together); rows absent from this call lose their dwell timers, keep cooldowns
      double fs[kMaxZonesPerStream], la[kMaxZonesPerStream];
      for (int c = 0; c < C; ++c) {
        fs[c] = kNaN;
        la[c] = 0.0;
        if (src >= 0) {
          la[c] = a.sin.last_alert[st0 + static_cast<size_t>(c) * cap + src];
          if (active) fs[c] = a.sin.first_seen[st0 + static_cast<size_t>(c) * cap + src];
        }
      }
      if (active) {
        cx = __float2int_rz(__fdiv_rn(__fadd_rn(box.x, box.z), 2.0f));
        cy = __float2int_rz(__fdiv_rn(__fadd_rn(box.y, box.w), 2.0f));
        for (int z = 0; z < nz; ++z) {
          const int p0 = s_zoff[z], k = s_zoff[z + 1] - p0, c = s_zcol[z];
          if (point_in_polygon(poly + p0, k, cx, cy) >= 0) {
            if (fs[c] != fs[c]) fs[c] = now;  // not in the zone before: start the dwell timer
            const double dwell = now - fs[c];
            if (dwell >= s_zdwell[z] && now - la[c] >= s_zcool[z]) {
              fired |= 1ull << z;
              dwell_of[z] = dwell;
              la[c] = now;
            }
          } else {
            fs[c] = kNaN;
          }
        }
      }
      for (int c = 0; c < C; ++c) {
        a.sout.first_seen[st0 + static_cast<size_t>(c) * cap + r] = fs[c];
        a.sout.last_alert[st0 + static_cast<size_t>(c) * cap + r] = la[c];
      }
    }
    // rank this round's events in (row, zone) order
    int tot;
    const int nfired = __popcll(fired);
    // exclusive prefix of per-thread event counts: warp shuffle scan + block scan of warp sums
    int incl = nfired;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(rtm::kFull, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int nwarp = kZoneThreads / 32;
      const int v = lane < nwarp ? s_scan[lane] : 0;
      int inc2 = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(rtm::kFull, inc2, d);
        if (lane >= d) inc2 += o;
      }
      if (lane < nwarp) s_scan[lane] = inc2 - v;
      if (lane == 31) s_scan[32] = inc2;
    }
    __syncthreads();
    int pos = ev_base + s_scan[warp] + incl - nfired;
    tot = s_scan[32];
    __syncthreads();
    while (fired) {
      const int z = __ffsll(static_cast<long long>(fired)) - 1;
      fired &= fired - 1;
      if (pos < a.event_stride) {
        rtm_zone_event e;
        e.stream = b;
        e.frame_id = a.frame_id;
        e.track_id = a.trk.track_id[row0 + r];
        e.zone = z;
        e.class_id = a.trk.class_id[row0 + r];
        e.cx = cx;
        e.cy = cy;
        e.row = r;
        e.dwell = dwell_of[z];
        e.now = now;
        e.xyxy[0] = box.x;
        e.xyxy[1] = box.y;
        e.xyxy[2] = box.z;
        e.xyxy[3] = box.w;
        ev_out[pos] = e;
      } else {
        overflow = true;
      }
      ++pos;
    }
    ev_base += tot;
  }
  if (overflow && a.status) atomicOr(&a.status[b], RTM_STATUS_EVENT_OVERFLOW);
  if (tid == 0) a.event_count[b] = min(ev_base, a.event_stride);
}

}  // namespace rtm
