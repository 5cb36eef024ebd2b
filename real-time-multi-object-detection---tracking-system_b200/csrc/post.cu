// rtm_post_backbone_step: the whole post-backbone step of B streams in two launches.
//
//   1. decode_tma / decode_scan (nms.cu)  head scan, candidate lists            HBM-bound
//   2. post_kernel (this file)            one CTA per stream: NMS -> tracker step -> zone step
//
// The three per-stream stages are small and strictly ordered; as separate kernels each paid a
// launch / drain overhead that was larger than its work.  Fused, the detections of a stream stay
// with the CTA that produced them and the step costs one dependent launch instead of three.  The
// stage bodies are the same device functions the stand-alone kernels run (nms_body.cuh,
// track_body.cuh, zone_body.cuh), so the fused step is bit-identical to rtm_decode_nms +
// rtm_track_step_ex + rtm_zone_step.  What does not depend on this frame's detections (head of the
// track table, zone table, polygons) is prefetched into shared memory behind the NMS.
//
// Consecutive steps overlap: the post kernel releases its dependents at once, and the next scan is
// a programmatic dependent launch (single stream), or runs on a stream of the library's own behind
// a caller-supplied "heads ready" event (rtm_step_io.scan_async) so that scans are back to back.
#include <stdlib.h>

#include <unordered_map>

#include "nms_body.cuh"
#include "track_body.cuh"
#include "zone_body.cuh"

namespace {

constexpr int kPostThreads = 512;
#ifndef RTM_POST_CTAS_PER_SM
#define RTM_POST_CTAS_PER_SM 1  // register budget: 2 = at most 64 registers
#endif

struct PostArgs {
  rtm::Workspace ws;
  rtm_nms_params prm;
  float iou_gate;
  rtm::NmsOut out;
  rtm::TrackArgs trk;
  rtm::ZoneArgs zone;
  int has_zones;
  int work_bytes;  // shared memory the three stages alias; the prefetch areas follow it
};

template <bool WITH_OPTIMAL>
__global__ void __launch_bounds__(kPostThreads, RTM_POST_CTAS_PER_SM) post_kernel(const __grid_constant__ PostArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_keep[rtm::kMaxDetCap];
  __shared__ int s_scan[33];
  // the inputs of the later stages that do not depend on this frame's detections are fetched
  // first, behind the NMS: the head of the track table, the zone table and its polygons
  rtm::TrackPrefetch* tpf = reinterpret_cast<rtm::TrackPrefetch*>(smem_raw + a.work_bytes);
  rtm::ZonePrefetch* zpf = reinterpret_cast<rtm::ZonePrefetch*>(tpf + 1);
  const int b = blockIdx.x;
  // Programmatic dependent launch: the next step's head scan (launched with programmatic stream
  // serialisation) may start now - it fills another slot of the candidate ring and touches nothing
  // this kernel reads or writes.  This kernel itself is an ordinary launch.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  RTM_TL(0);
  auto prefetch = [&]() {
    if (a.has_zones) rtm::zone_prefetch<kPostThreads>(a.zone, b, zpf);
    rtm::track_prefetch<kPostThreads>(a.trk, b, tpf);
  };
  rtm::nms_stream<kPostThreads>(a.ws, a.prm, a.iou_gate, a.out, b, smem_raw, s_keep, s_scan, prefetch);  // ends with a barrier
  RTM_TL(10);
  rtm::track_stream<kPostThreads, WITH_OPTIMAL>(a.trk, b, smem_raw, tpf);
  __syncthreads();
  RTM_TL(20);
  if (a.has_zones) rtm::zone_stream<kPostThreads>(a.zone, b, smem_raw, s_scan, zpf);
  RTM_TL(30);
}

#ifdef RTM_TIMELINE
}  // namespace
extern "C" int rtm_debug_timeline(void* device_buffer) {  // (B, 32) u64, or null to stop
  unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
  RTM_CUDA(cudaMemcpyToSymbol(rtm::g_timeline, &p, sizeof(p)));
  return RTM_OK;
}
namespace {
#endif


// scan_async: per workspace, the stream the scans go to and the events that order it with the caller's stream
struct ScanCtx {
  cudaStream_t stream = nullptr;
  cudaEvent_t scanned[rtm::kCandSlots] = {};   // slot's candidate list is complete
  cudaEvent_t consumed[rtm::kCandSlots] = {};  // slot's post kernel is done (recorded on the caller's stream)
  bool consumed_valid[rtm::kCandSlots] = {};
  int covered = 0;  // scans to come (on `stream`) whose slots are already known to be free: see the waits below
};

std::unordered_map<const void*, ScanCtx>& scan_table() {
  static std::unordered_map<const void*, ScanCtx> table;
  return table;
}

int scan_ctx(const void* workspace, ScanCtx** out) {
  ScanCtx& c = scan_table()[workspace];
  if (!c.stream) {
    RTM_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (int i = 0; i < rtm::kCandSlots; ++i) {
      RTM_CUDA(cudaEventCreateWithFlags(&c.scanned[i], cudaEventDisableTiming));
      RTM_CUDA(cudaEventCreateWithFlags(&c.consumed[i], cudaEventDisableTiming));
    }
  }
  *out = &c;
  return RTM_OK;
}

#ifdef RTM_PROBES
// tools/probe_overlap.py: what the orderings around the scan cost (results are NOT valid with any bit set)
int probe_bits() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTM_PROBE_BITS");
    v = e ? atoi(e) : 0;
  }
  return v;
}
#endif

bool fuse_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTM_FUSE_POST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

extern "C" int rtm_post_backbone_step(const rtm_step_io* io, const rtm_nms_params* params,
                                      rtm_cuda_stream stream) {
  RTM_REQUIRE(io && params, "rtm_post_backbone_step: null argument");
  RTM_REQUIRE(io->table_in && io->table_out, "rtm_post_backbone_step: null track table");
  const int B = io->table_in->num_streams;
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  const bool optimal = io->assignment == RTM_ASSIGN_OPTIMAL;
  RTM_REQUIRE(io->assignment == RTM_ASSIGN_GREEDY || optimal, "rtm_post_backbone_step: unknown assignment mode %d", io->assignment);
  const size_t track_smem = rtm::track_smem_bytes(io->det_stride, io->table_in->capacity, optimal);
  size_t work = rtm::kNmsSmemBytes;
  if (track_smem > work) work = track_smem;
  if (io->zones && rtm::zone_smem_bytes(io->event_stride) > work) work = rtm::zone_smem_bytes(io->event_stride);
  work = (work + 127) / 128 * 128;
  const size_t smem = work + sizeof(rtm::TrackPrefetch) + sizeof(rtm::ZonePrefetch);

  if (!fuse_enabled() || smem > 200 * 1024) {
    // unfused fallback: the three stand-alone entry points back to back, all on the caller's stream
    if (io->scan_async && io->heads_ready_event)
      RTM_CUDA(cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(io->heads_ready_event), 0));
    int rc = rtm_decode_nms(io->head_p3, io->head_p4, io->head_p5, io->head_dtype, B, io->img_h, io->img_w, params,
                            io->scale, io->det_xyxy, io->det_conf, io->det_cls, io->det_anchor, io->det_keep,
                            io->det_count, io->det_stride, io->status, io->workspace, io->workspace_bytes, stream);
    if (rc) return rc;
    const rtm_track_options opt{io->track_thresh, io->match_thresh, io->track_buffer, io->assignment, io->kalman_in,
                                io->kalman_out, io->cost_limit};
    rc = rtm_track_step_ex(io->table_in, io->table_out, io->det_xyxy, io->det_conf, io->det_cls, io->det_count,
                           io->det_stride, &opt, io->det_track_id, io->det_kind, io->src_row, io->status, stream);
    if (rc) return rc;
    if (!io->zones) return RTM_OK;
    return rtm_zone_step(io->zones, io->table_out, io->src_row, io->state_in, io->state_out, io->now,
                         io->now_per_stream, io->frame_id, io->events, io->event_stride, io->event_count,
                         io->status, stream);
  }

  // ---- argument checks of the three stages ----
  RTM_REQUIRE(io->det_xyxy && io->det_conf && io->det_cls && io->det_count && io->workspace, "rtm_post_backbone_step: null detection buffers");
  RTM_REQUIRE(params->max_det > 0 && params->max_det <= rtm::kMaxDetCap && io->det_stride >= params->max_det,
              "rtm_post_backbone_step: max_det %d / det_stride %d out of range", params->max_det, io->det_stride);
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(io->det_xyxy) & 15) == 0, "det_xyxy must be 16-byte aligned");
  RTM_REQUIRE(io->table_in->num_streams == io->table_out->num_streams && io->table_in->capacity == io->table_out->capacity &&
                  io->table_in->capacity > 0 && B > 0, "rtm_post_backbone_step: bad track tables");
  RTM_REQUIRE(io->table_in->xyxy != io->table_out->xyxy, "rtm_post_backbone_step: table_in and table_out must be distinct");
  RTM_REQUIRE((io->kalman_in == nullptr) == (io->kalman_out == nullptr), "rtm_post_backbone_step: kalman_in / kalman_out go together");
  if (io->kalman_in)
    RTM_REQUIRE(io->kalman_in->mean && io->kalman_in->cov && io->kalman_out->mean && io->kalman_out->cov &&
                    io->kalman_in->mean != io->kalman_out->mean, "rtm_post_backbone_step: bad Kalman state arrays");
  if (io->zones) {
    RTM_REQUIRE(io->state_in && io->state_out && io->events && io->event_count && io->event_stride > 0 && io->src_row,
                "rtm_post_backbone_step: incomplete zone arguments");
    RTM_REQUIRE(io->zones->num_streams == B && io->zones->num_columns > 0, "rtm_post_backbone_step: zone set shape");
    RTM_REQUIRE(io->state_in->first_seen != io->state_out->first_seen, "rtm_post_backbone_step: zone state must ping-pong");
  }

  PostArgs a;
  ScanCtx* ctx = nullptr;
  int rc;
  if (io->scan_async) {
    // the scan goes to the library's own stream, behind "heads ready" and behind the post kernel that last
    // read the slot it is about to fill; the caller's stream waits for it before the post kernel
    rc = scan_ctx(io->workspace, &ctx);
    if (rc) return rc;
    const int slot = rtm::next_scan_slot(io->workspace);
    // (waits that are already satisfied are not enqueued: they would sit between consecutive scans)
    if (io->heads_ready_event && cudaEventQuery(static_cast<cudaEvent_t>(io->heads_ready_event)) != cudaSuccess)
      RTM_CUDA(cudaStreamWaitEvent(ctx->stream, static_cast<cudaEvent_t>(io->heads_ready_event), 0));
#ifdef RTM_PROBES
    if (!(probe_bits() & 2))  // timing probe only: slot reuse unordered
#endif
    if (ctx->covered == 0) {
      // this scan's slot and the ones after it up to the next multiple of kSlotWaitEvery: their last readers
      // (post kernels of at least kCandSlots - kSlotWaitEvery + 1 steps ago) must be done before they are refilled
      const int group_end = (slot / rtm::kSlotWaitEvery + 1) * rtm::kSlotWaitEvery;
      for (int sl = slot; sl < group_end; ++sl)
        if (ctx->consumed_valid[sl] && cudaEventQuery(ctx->consumed[sl]) != cudaSuccess)
          RTM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->consumed[sl], 0));
      ctx->covered = group_end - slot;
    }
    --ctx->covered;
    (void)cudaGetLastError();  // cudaEventQuery reports "not ready" through the error state
    rc = rtm::launch_decode_stage(io->head_p3, io->head_p4, io->head_p5, io->head_dtype, B, io->img_h, io->img_w, params,
                                  io->workspace, io->workspace_bytes, &a.ws, ctx->stream);
    if (rc) return rc;
#ifdef RTM_PROBES
    if (!(probe_bits() & 1))  // timing probe only: the post kernel races with its scan
#endif
    {
      RTM_CUDA(cudaEventRecord(ctx->scanned[a.ws.slot], ctx->stream));
      RTM_CUDA(cudaStreamWaitEvent(s, ctx->scanned[a.ws.slot], 0));
    }
  } else {
    rc = rtm::launch_decode_stage(io->head_p3, io->head_p4, io->head_p5, io->head_dtype, B, io->img_h, io->img_w, params,
                                  io->workspace, io->workspace_bytes, &a.ws, s);
    if (rc) return rc;
    // a workspace that has been stepped with scan_async before keeps its slot bookkeeping up to date
    const auto it = scan_table().find(io->workspace);
    if (it != scan_table().end()) {
      ctx = &it->second;
      ctx->covered = 0;  // this scan is not on the scan stream: the next asynchronous one orders itself afresh
    }
  }
  a.prm = *params;
  a.iou_gate = rtm::iou_gate_for(params->iou_thres);
  a.out = rtm::NmsOut{io->scale, io->det_xyxy, io->det_conf, io->det_cls, io->det_anchor, io->det_keep, io->det_count,
                      io->det_stride, io->status};
  a.trk = rtm::TrackArgs{*io->table_in, *io->table_out, io->det_xyxy, io->det_conf, io->det_cls, io->det_count,
                         io->det_stride, io->track_thresh, io->match_thresh, io->track_buffer, io->det_track_id,
                         io->det_kind, io->src_row, io->status,
                         nullptr, nullptr, nullptr, nullptr, io->assignment, io->cost_limit};
  if (io->kalman_in) {
    a.trk.kf_mean_in = io->kalman_in->mean;
    a.trk.kf_cov_in = io->kalman_in->cov;
    a.trk.kf_mean_out = io->kalman_out->mean;
    a.trk.kf_cov_out = io->kalman_out->cov;
  }
  a.has_zones = io->zones != nullptr;
  if (a.has_zones)
    a.zone = rtm::ZoneArgs{*io->zones, *io->table_out, io->src_row, *io->state_in, *io->state_out, io->now,
                           io->now_per_stream, io->frame_id, io->events, io->event_stride, io->event_count,
                           io->status};
  a.work_bytes = static_cast<int>(work);
  static size_t configured = 0;
  if (smem > configured) {
    RTM_CUDA(cudaFuncSetAttribute(post_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    RTM_CUDA(cudaFuncSetAttribute(post_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  {
    rtm::ProfileScope prof(RTM_K_POST, s);
    // An ordinary launch on purpose.  Launching this kernel as a programmatic dependent of its own
    // step's scan as well was measured slower (56.7 vs 46.0 us per step: its CTAs need whole SMs and
    // hold them while they wait) and would need the scan to wait for the previous post kernel.
#ifdef RTM_PROBES
    if (probe_bits() & 4)  // timing probe only: scans alone (the NMS stage would have re-armed the slot's ticket counter)
      RTM_CUDA(cudaMemsetAsync(a.ws.tile_counter, 0, sizeof(int), s));
    else
#endif
    if (optimal) post_kernel<true><<<B, kPostThreads, smem, s>>>(a);
    else post_kernel<false><<<B, kPostThreads, smem, s>>>(a);
  }
  RTM_LAUNCH_CHECK("post_kernel");
  if (ctx) {
    RTM_CUDA(cudaEventRecord(ctx->consumed[a.ws.slot], s));
    ctx->consumed_valid[a.ws.slot] = true;
  }
  return RTM_OK;
}
