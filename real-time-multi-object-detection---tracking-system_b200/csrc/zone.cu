// Z1..Z3: one zone-engine step for B independent streams (rtm_zone_step).
//
// Semantics follow ZoneEventEngine.process (src/events/zone_engine.py:82-132):
//   centroid = int((x1+x2)/2), int((y1+y2)/2) in float32, truncated toward zero (:90-91);
//   inside   = cv2.pointPolygonTest(poly_int32, (cx, cy), False) >= 0 (:94) - the integer
//              crossing test below reproduces OpenCV's routine, boundary counts as inside;
//   inside  -> first_seen is set once, dwell = now - first_seen; an event fires when
//              dwell >= dwell_time_sec and now - last_alert >= cooldown_sec (:96-121);
//   outside -> the dwell timer of that zone name is dropped (:122-125);
//   tracks not passed to the call lose all dwell timers, cooldowns persist (:128-130).
//
// One CTA per stream.  The zone table and polygons of the stream are staged in shared memory; one
// thread per (track row, state column) walks the column's zones IN ORDER (zones sharing a name
// share a state column, so their order matters and is kept; different columns never interact).
// Fired (row, zone) pairs are ranked with a block scan so the event list comes out in the
// reference's (track order, zone order).
#include <stdlib.h>

#include "zone_body.cuh"

namespace {

using rtm::ZoneArgs;

// 256 threads per stream, or 1024 when the tables are large (crowds: the step is a chain of dependent loads per
// (row, column) pair, so the pairs a thread walks one after the other set the time - ncu: 12 % of the warp slots
// active at 256 threads and 128 streams)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) zone_step_kernel(const ZoneArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[33];
  __shared__ rtm::ZonePrefetch zp;
  rtm::zone_prefetch<THREADS>(a, blockIdx.x, &zp);
  __syncthreads();
  rtm::zone_stream<THREADS>(a, blockIdx.x, smem_raw, s_scan, &zp);
}

template <int THREADS>
int launch_zone(const ZoneArgs& a, size_t smem, cudaStream_t stream) {
  if (smem > 24 * 1024)
    if (int rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(zone_step_kernel<THREADS>), smem)) return rc;
  {
    rtm::ProfileScope prof(RTM_K_ZONE, stream);
    zone_step_kernel<THREADS><<<a.trk.num_streams, THREADS, smem, stream>>>(a);
  }
  RTM_LAUNCH_CHECK("zone_step_kernel");
  return RTM_OK;
}

}  // namespace

extern "C" int rtm_zone_step(const rtm_zone_set* zones, const rtm_track_table* tracks,
                             const int32_t* src_row, const rtm_zone_state* state_in,
                             const rtm_zone_state* state_out, double now, const double* now_per_stream,
                             int32_t frame_id, rtm_zone_event* events, int32_t event_stride,
                             int32_t* event_count, int32_t* status, rtm_cuda_stream stream) {
  RTM_REQUIRE(zones && tracks && state_in && state_out, "rtm_zone_step: null argument");
  RTM_REQUIRE(zones->num_streams == tracks->num_streams, "rtm_zone_step: zone set has %d streams, track table %d",
              zones->num_streams, tracks->num_streams);
  RTM_REQUIRE(zones->num_columns > 0 && zones->zone_offsets && zones->poly_offsets && zones->column,
              "rtm_zone_step: incomplete zone set");
  RTM_REQUIRE(events && event_count && event_stride > 0, "rtm_zone_step: null event buffers");
  RTM_REQUIRE(!(src_row && state_in->first_seen == state_out->first_seen),
              "rtm_zone_step: in-place state needs src_row == NULL (identity)");
  ZoneArgs a{*zones, *tracks, src_row, *state_in, *state_out, now, now_per_stream, frame_id,
             events, event_stride, event_count, status};
  const size_t smem = rtm::zone_smem_bytes(event_stride);
  RTM_REQUIRE(smem + sizeof(rtm::ZonePrefetch) <= 226 * 1024, "rtm_zone_step: %zu B of shared memory needed", smem);
  static const int forced = getenv("RTM_ZONE_THREADS") ? atoi(getenv("RTM_ZONE_THREADS")) : 0;
  const bool wide = forced ? forced == 1024 : static_cast<long long>(tracks->capacity) * zones->num_columns >= 8192;
  return wide ? launch_zone<1024>(a, smem, static_cast<cudaStream_t>(stream)) : launch_zone<256>(a, smem, static_cast<cudaStream_t>(stream));
}
