// T1..T4: one tracker step for B independent streams (rtm_track_step).
//
// Semantics follow the reference's _ByteTrackCore (src/tracking/tracker.py:58-194, greedy
// assignment branch): two-stage IoU association against the stored boxes, births from
// unmatched high-score detections, ageing, pruning.  One CTA per stream.
//
//   * detections of the stream are staged in shared memory (box, area, high/low lists);
//   * a warp takes a track row at a time: its lanes split the detection tile, each keeping the
//     first arg-max of its columns in registers, then combine by shuffle - the T x N IoU matrix
//     is never materialised (tracker.py:150-161 builds it; only its row arg-max is ever read);
//   * the row-order greedy loop of tracker.py:186-191 is evaluated in its order-free form:
//     an admissible row (max IoU >= thresh) bids for its arg-max column with atomicMin(row);
//     the smallest row wins, every other bidder stays unmatched (no second choice);
//   * births / survivors are placed with order-preserving block scans, so the output table
//     is in creation order exactly like the reference's list.
//
// Arithmetic is float32 with explicitly rounded operations (no FMA contraction), matching
// NumPy's evaluation of tracker.py:153-161 bit for bit for finite boxes.
#include "track_body.cuh"

namespace {

using rtm::TrackArgs;
using rtm::track_smem_bytes;

template <int THREADS>
__global__ void __launch_bounds__(THREADS) track_step_kernel(const TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ rtm::TrackPrefetch pf;
  rtm::track_prefetch<THREADS>(a, blockIdx.x, &pf);
  __syncthreads();
  rtm::track_stream<THREADS, true>(a, blockIdx.x, smem_raw, &pf);
}

template <int THREADS>
int launch_track(const TrackArgs& a, cudaStream_t stream) {
  const size_t smem = track_smem_bytes(a.det_stride, a.tin.capacity, a.assignment == RTM_ASSIGN_OPTIMAL);
  RTM_REQUIRE(smem + sizeof(rtm::TrackPrefetch) <= 226 * 1024, "rtm_track_step: det_stride %d / capacity %d need %zu B of shared memory (> 227 KB)",
              a.det_stride, a.tin.capacity, smem);
  if (smem + sizeof(rtm::TrackPrefetch) > 40 * 1024)  // static + dynamic beyond the default limit
    if (int rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(track_step_kernel<THREADS>), smem)) return rc;
  {
    rtm::ProfileScope prof(RTM_K_TRACK, stream);
    track_step_kernel<THREADS><<<a.tin.num_streams, THREADS, smem, stream>>>(a);
  }
  RTM_LAUNCH_CHECK("track_step_kernel");
  return RTM_OK;
}

}  // namespace

extern "C" int rtm_track_step_ex(const rtm_track_table* table_in, const rtm_track_table* table_out,
                                 const float* det_xyxy, const float* det_conf, const int32_t* det_cls,
                                 const int32_t* det_count, int32_t det_stride, const rtm_track_options* opt,
                                 int32_t* det_track_id, int32_t* det_kind, int32_t* src_row, int32_t* status,
                                 rtm_cuda_stream stream) {
  RTM_REQUIRE(table_in && table_out && opt, "rtm_track_step: null table / options");
  RTM_REQUIRE(table_in->num_streams == table_out->num_streams && table_in->capacity == table_out->capacity,
              "rtm_track_step: table_in / table_out shapes differ");
  RTM_REQUIRE(table_in->num_streams > 0 && table_in->capacity > 0, "rtm_track_step: empty table");
  RTM_REQUIRE(table_in->xyxy != table_out->xyxy, "rtm_track_step: table_in and table_out must be distinct");
  RTM_REQUIRE(det_xyxy && det_conf && det_cls && det_count && det_stride > 0, "rtm_track_step: null detections");
  RTM_REQUIRE((opt->kalman_in == nullptr) == (opt->kalman_out == nullptr), "rtm_track_step: kalman_in / kalman_out go together");
  RTM_REQUIRE(opt->assignment == RTM_ASSIGN_GREEDY || opt->assignment == RTM_ASSIGN_OPTIMAL, "rtm_track_step: unknown assignment mode %d",
              opt->assignment);
  TrackArgs a{*table_in, *table_out, det_xyxy, det_conf, det_cls, det_count, det_stride,
              opt->track_thresh, opt->match_thresh, opt->track_buffer, det_track_id, det_kind, src_row, status,
              nullptr, nullptr, nullptr, nullptr, opt->assignment, opt->cost_limit, nullptr, 0};
  if (opt->assign_scratch && opt->assign_scratch_bytes) {
    RTM_REQUIRE((reinterpret_cast<uintptr_t>(opt->assign_scratch) & 15) == 0, "rtm_track_step: assign_scratch must be 16-byte aligned");
    a.assign_scratch = static_cast<unsigned char*>(opt->assign_scratch);
    a.assign_scratch_per_stream = opt->assign_scratch_bytes / table_in->num_streams / 16 * 16;
  }
  if (opt->kalman_in) {
    RTM_REQUIRE(opt->kalman_in->mean && opt->kalman_in->cov && opt->kalman_out->mean && opt->kalman_out->cov,
                "rtm_track_step: null Kalman state arrays");
    RTM_REQUIRE(opt->kalman_in->mean != opt->kalman_out->mean, "rtm_track_step: Kalman state must ping-pong like the tables");
    a.kf_mean_in = opt->kalman_in->mean;
    a.kf_cov_in = opt->kalman_in->cov;
    a.kf_mean_out = opt->kalman_out->mean;
    a.kf_cov_out = opt->kalman_out->cov;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (det_stride > 256 || table_in->capacity > 2048) return launch_track<1024>(a, s);
  return launch_track<256>(a, s);
}

extern "C" int rtm_track_step(const rtm_track_table* table_in, const rtm_track_table* table_out,
                              const float* det_xyxy, const float* det_conf, const int32_t* det_cls,
                              const int32_t* det_count, int32_t det_stride, float track_thresh,
                              float match_thresh, int32_t track_buffer, int32_t* det_track_id,
                              int32_t* det_kind, int32_t* src_row, int32_t* status,
                              rtm_cuda_stream stream) {
  const rtm_track_options opt{track_thresh, match_thresh, track_buffer, RTM_ASSIGN_GREEDY, nullptr, nullptr, 0.0, nullptr, 0};
  return rtm_track_step_ex(table_in, table_out, det_xyxy, det_conf, det_cls, det_count, det_stride, &opt, det_track_id,
                           det_kind, src_row, status, stream);
}

extern "C" size_t rtm_assign_scratch_bytes(int32_t num_streams, int32_t capacity, int32_t det_stride, int32_t max_pairs) {
  if (num_streams <= 0 || capacity <= 0 || det_stride <= 0 || max_pairs < 0) return 0;
  const size_t per = (rtm::assign_global_fixed_bytes(capacity, det_stride) + static_cast<size_t>(max_pairs) * 6 + 31) / 16 * 16;
  return per * num_streams;
}
