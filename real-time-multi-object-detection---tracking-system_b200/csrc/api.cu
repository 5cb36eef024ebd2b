// Library-level entry points: version, error string, device info, and the fused
// post-backbone step (device-resident and host-fed variants).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "rtm_common.cuh"

namespace rtm {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int& c = cached[dev & 63];
  if (!c) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c;
}

int ensure_dynamic_smem(const void* func, size_t bytes) {
  static std::mutex m;
  static std::map<std::pair<const void*, int>, size_t> configured;  // per (kernel, device)
  int dev = 0;
  RTM_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(m);
  size_t& have = configured[std::make_pair(func, dev)];
  if (bytes > have) {
    RTM_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    have = bytes;
  }
  return RTM_OK;
}

bool g_profile_on = false;

// RTM_PDL=0 turns programmatic dependent launch of the step's two kernels off
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RTM_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

namespace {
struct ProfileRecord {
  int kind;
  cudaEvent_t start, stop;
};
std::vector<ProfileRecord> g_records;
std::vector<cudaEvent_t> g_free_events;
std::mutex g_profile_mutex;

cudaEvent_t take_event() {
  if (!g_free_events.empty()) {
    cudaEvent_t e = g_free_events.back();
    g_free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void profile_begin(int kind, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  ProfileRecord r{kind, take_event(), take_event()};
  cudaEventRecord(r.start, s);
  g_records.push_back(r);
}

void profile_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  if (!g_records.empty()) cudaEventRecord(g_records.back().stop, s);
}

}  // namespace rtm

extern "C" int rtm_profile_enable(int32_t on) {
  rtm::g_profile_on = on != 0;
  return RTM_OK;
}

extern "C" int rtm_profile_read(double* ms_sum, int32_t* launches) {
  RTM_REQUIRE(ms_sum && launches, "rtm_profile_read: null output");
  std::lock_guard<std::mutex> lock(rtm::g_profile_mutex);
  for (int k = 0; k < RTM_K_COUNT; ++k) {
    ms_sum[k] = 0.0;
    launches[k] = 0;
  }
  for (const auto& r : rtm::g_records) {
    RTM_CUDA(cudaEventSynchronize(r.stop));
    float ms = 0.f;
    RTM_CUDA(cudaEventElapsedTime(&ms, r.start, r.stop));
    if (r.kind >= 0 && r.kind < RTM_K_COUNT) {
      ms_sum[r.kind] += ms;
      launches[r.kind] += 1;
    }
    rtm::g_free_events.push_back(r.start);
    rtm::g_free_events.push_back(r.stop);
  }
  rtm::g_records.clear();
  return RTM_OK;
}

extern "C" int rtm_version(void) { return RTM_VERSION; }

extern "C" const char* rtm_last_error(void) { return rtm::g_error; }

extern "C" int rtm_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0, sms = 0, major = 0, minor = 0;
  RTM_CUDA(cudaGetDevice(&dev));
  RTM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RTM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  RTM_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  if (major != 10) {
    rtm::set_error("librtmodt_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return RTM_ERR_UNSUPPORTED;
  }
  return RTM_OK;
}

namespace {

size_t elem_size(int dtype) { return dtype == RTM_F32 ? 4 : 2; }

}  // namespace

extern "C" int rtm_post_backbone_step_host(const rtm_step_io* io, const rtm_step_host_io* h,
                                           const rtm_nms_params* params, rtm_cuda_stream stream) {
  RTM_REQUIRE(io && h && params, "rtm_post_backbone_step_host: null argument");
  RTM_REQUIRE(h->host_head_p3 && h->host_head_p4 && h->host_head_p5, "rtm_post_backbone_step_host: null host head");
  RTM_REQUIRE(io->table_in, "rtm_post_backbone_step_host: null track table");
  RTM_REQUIRE(io->img_h > 0 && io->img_w > 0 && io->img_h % 32 == 0 && io->img_w % 32 == 0,
              "rtm_post_backbone_step_host: bad image size");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t B = io->table_in->num_streams;
  const size_t ch = 64 + params->num_classes, es = elem_size(io->head_dtype);
  const void* src[3] = {h->host_head_p3, h->host_head_p4, h->host_head_p5};
  void* dst[3] = {const_cast<void*>(io->head_p3), const_cast<void*>(io->head_p4), const_cast<void*>(io->head_p5)};
  const int strides[3] = {8, 16, 32};
  size_t bytes[3];
  for (int l = 0; l < 3; ++l) bytes[l] = B * ch * (static_cast<size_t>(io->img_h / strides[l]) * (io->img_w / strides[l])) * es;
  const char *s0 = static_cast<const char*>(src[0]), *d0 = static_cast<const char*>(dst[0]);
  if (src[1] == s0 + bytes[0] && src[2] == s0 + bytes[0] + bytes[1] && dst[1] == d0 + bytes[0] && dst[2] == d0 + bytes[0] + bytes[1]) {
    // the three levels are back to back on both sides: one transfer (a few percent more PCIe throughput)
    RTM_CUDA(cudaMemcpyAsync(dst[0], src[0], bytes[0] + bytes[1] + bytes[2], cudaMemcpyHostToDevice, s));
  } else {
    for (int l = 0; l < 3; ++l) RTM_CUDA(cudaMemcpyAsync(dst[l], src[l], bytes[l], cudaMemcpyHostToDevice, s));
  }
  if (h->wait_event) RTM_CUDA(cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(h->wait_event), 0));
  int rc = rtm_post_backbone_step(io, params, stream);
  if (rc) return rc;
  const size_t D = static_cast<size_t>(io->det_stride);
  if (io->zones && h->host_events && h->host_event_count) {
    RTM_CUDA(cudaMemcpyAsync(h->host_event_count, io->event_count, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    const size_t hs = h->host_event_stride > 0 && h->host_event_stride < io->event_stride ? h->host_event_stride : io->event_stride;
    if (hs == static_cast<size_t>(io->event_stride))
      RTM_CUDA(cudaMemcpyAsync(h->host_events, io->events, B * hs * sizeof(rtm_zone_event), cudaMemcpyDeviceToHost, s));
    else  // the head of every stream's slab only
      RTM_CUDA(cudaMemcpy2DAsync(h->host_events, hs * sizeof(rtm_zone_event), io->events, io->event_stride * sizeof(rtm_zone_event),
                                 hs * sizeof(rtm_zone_event), B, cudaMemcpyDeviceToHost, s));
  }
  if (h->host_det_count) RTM_CUDA(cudaMemcpyAsync(h->host_det_count, io->det_count, B * 4, cudaMemcpyDeviceToHost, s));
  if (h->host_det_xyxy) RTM_CUDA(cudaMemcpyAsync(h->host_det_xyxy, io->det_xyxy, B * D * 16, cudaMemcpyDeviceToHost, s));
  if (h->host_det_conf) RTM_CUDA(cudaMemcpyAsync(h->host_det_conf, io->det_conf, B * D * 4, cudaMemcpyDeviceToHost, s));
  if (h->host_det_cls) RTM_CUDA(cudaMemcpyAsync(h->host_det_cls, io->det_cls, B * D * 4, cudaMemcpyDeviceToHost, s));
  if (h->host_det_track_id && io->det_track_id)
    RTM_CUDA(cudaMemcpyAsync(h->host_det_track_id, io->det_track_id, B * D * 4, cudaMemcpyDeviceToHost, s));
  if (h->host_status && io->status)
    RTM_CUDA(cudaMemcpyAsync(h->host_status, io->status, B * 4, cudaMemcpyDeviceToHost, s));
  if (h->done_event) RTM_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(h->done_event), s));
  return RTM_OK;
}
