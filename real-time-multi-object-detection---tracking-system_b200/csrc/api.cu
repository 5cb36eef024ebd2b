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
  if (h->copy_wait_event) RTM_CUDA(cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(h->copy_wait_event), 0));
  const char *s0 = static_cast<const char*>(src[0]), *d0 = static_cast<const char*>(dst[0]);
  if (src[1] == s0 + bytes[0] && src[2] == s0 + bytes[0] + bytes[1] && dst[1] == d0 + bytes[0] && dst[2] == d0 + bytes[0] + bytes[1]) {
    // the three levels are back to back on both sides: one transfer (a few percent more PCIe throughput)
    RTM_CUDA(cudaMemcpyAsync(dst[0], src[0], bytes[0] + bytes[1] + bytes[2], cudaMemcpyHostToDevice, s));
  } else {
    for (int l = 0; l < 3; ++l) RTM_CUDA(cudaMemcpyAsync(dst[l], src[l], bytes[l], cudaMemcpyHostToDevice, s));
  }
  if (h->copy_done_event) RTM_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(h->copy_done_event), s));
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


// ---------------------------------------------------------------------------------------
// checkpoint / resume
// ---------------------------------------------------------------------------------------
namespace {

struct StateHeader {  // 64 bytes
  char magic[8];      // "RTMSTATE"
  int32_t version, num_streams, capacity, num_columns, with_zone_state, with_kalman;
  int32_t reserved[8];
};
static_assert(sizeof(StateHeader) == 64, "state blob header");

struct StatePiece {
  void* dev;
  size_t bytes;
};

// the arrays of a state blob in the order they are stored
int state_pieces(const rtm_track_table* t, const rtm_zone_state* z, int num_columns, const rtm_kalman_state* k, StatePiece* out) {
  const size_t B = t->num_streams, rows = B * static_cast<size_t>(t->capacity);
  int n = 0;
  out[n++] = {t->count, B * 4};
  out[n++] = {t->next_id, B * 4};
  out[n++] = {t->track_id, rows * 4};
  out[n++] = {t->xyxy, rows * 16};
  out[n++] = {t->confidence, rows * 4};
  out[n++] = {t->class_id, rows * 4};
  out[n++] = {t->age, rows * 4};
  out[n++] = {t->time_since_update, rows * 4};
  if (z) {
    out[n++] = {z->first_seen, rows * num_columns * 8};
    out[n++] = {z->last_alert, rows * num_columns * 8};
  }
  if (k) {
    out[n++] = {k->mean, rows * 8 * 4};
    out[n++] = {k->cov, rows * 12 * 4};
  }
  return n;
}

int check_state_args(const rtm_track_table* t, const rtm_zone_state* z, int num_columns, const rtm_kalman_state* k) {
  RTM_REQUIRE(t && t->num_streams > 0 && t->capacity > 0, "rtm_state: bad track table");
  RTM_REQUIRE(t->count && t->next_id && t->track_id && t->xyxy && t->confidence && t->class_id && t->age && t->time_since_update,
              "rtm_state: null track table array");
  if (z) RTM_REQUIRE(z->first_seen && z->last_alert && num_columns > 0, "rtm_state: bad zone state");
  if (k) RTM_REQUIRE(k->mean && k->cov, "rtm_state: bad Kalman state");
  return RTM_OK;
}

}  // namespace

extern "C" size_t rtm_state_bytes(int32_t num_streams, int32_t capacity, int32_t num_columns, int32_t with_zone_state,
                                  int32_t with_kalman) {
  if (num_streams <= 0 || capacity <= 0) return 0;
  const size_t B = num_streams, rows = B * static_cast<size_t>(capacity);
  size_t n = sizeof(StateHeader) + B * 8 + rows * (4 + 16 + 4 + 4 + 4 + 4);
  if (with_zone_state) n += rows * static_cast<size_t>(num_columns > 0 ? num_columns : 0) * 16;
  if (with_kalman) n += rows * 20 * 4;
  return n;
}

extern "C" int rtm_state_export(const rtm_track_table* table, const rtm_zone_state* zone_state, int32_t num_columns,
                                const rtm_kalman_state* kalman, void* host_blob, size_t blob_bytes, rtm_cuda_stream stream) {
  int rc = check_state_args(table, zone_state, num_columns, kalman);
  if (rc) return rc;
  const size_t need = rtm_state_bytes(table->num_streams, table->capacity, num_columns, zone_state != nullptr, kalman != nullptr);
  RTM_REQUIRE(host_blob && blob_bytes >= need, "rtm_state_export: blob has %zu bytes, %zu needed", blob_bytes, need);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StateHeader h = {};
  memcpy(h.magic, "RTMSTATE", 8);
  h.version = 1;
  h.num_streams = table->num_streams;
  h.capacity = table->capacity;
  h.num_columns = zone_state ? num_columns : 0;
  h.with_zone_state = zone_state != nullptr;
  h.with_kalman = kalman != nullptr;
  memcpy(host_blob, &h, sizeof(h));
  StatePiece pieces[12];
  const int n = state_pieces(table, zone_state, num_columns, kalman, pieces);
  char* p = static_cast<char*>(host_blob) + sizeof(h);
  for (int i = 0; i < n; ++i) {
    RTM_CUDA(cudaMemcpyAsync(p, pieces[i].dev, pieces[i].bytes, cudaMemcpyDeviceToHost, s));
    p += pieces[i].bytes;
  }
  RTM_CUDA(cudaStreamSynchronize(s));
  return RTM_OK;
}

extern "C" int rtm_state_import(const rtm_track_table* table, const rtm_zone_state* zone_state, int32_t num_columns,
                                const rtm_kalman_state* kalman, const void* host_blob, size_t blob_bytes, rtm_cuda_stream stream) {
  int rc = check_state_args(table, zone_state, num_columns, kalman);
  if (rc) return rc;
  RTM_REQUIRE(host_blob && blob_bytes >= sizeof(StateHeader), "rtm_state_import: blob too small");
  StateHeader h;
  memcpy(&h, host_blob, sizeof(h));
  RTM_REQUIRE(memcmp(h.magic, "RTMSTATE", 8) == 0 && h.version == 1, "rtm_state_import: not a state blob (or another version)");
  RTM_REQUIRE(h.num_streams == table->num_streams && h.capacity == table->capacity,
              "rtm_state_import: blob holds %d streams x %d rows, the tables %d x %d", h.num_streams, h.capacity, table->num_streams,
              table->capacity);
  RTM_REQUIRE((h.with_zone_state != 0) == (zone_state != nullptr) && (!zone_state || h.num_columns == num_columns),
              "rtm_state_import: zone state of the blob (%d columns) and of the call (%d) differ", h.with_zone_state ? h.num_columns : 0,
              zone_state ? num_columns : 0);
  RTM_REQUIRE((h.with_kalman != 0) == (kalman != nullptr), "rtm_state_import: Kalman state present on one side only");
  const size_t need = rtm_state_bytes(h.num_streams, h.capacity, h.num_columns, h.with_zone_state, h.with_kalman);
  RTM_REQUIRE(blob_bytes >= need, "rtm_state_import: blob has %zu bytes, %zu needed", blob_bytes, need);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StatePiece pieces[12];
  const int n = state_pieces(table, zone_state, num_columns, kalman, pieces);
  const char* p = static_cast<const char*>(host_blob) + sizeof(h);
  for (int i = 0; i < n; ++i) {
    RTM_CUDA(cudaMemcpyAsync(pieces[i].dev, p, pieces[i].bytes, cudaMemcpyHostToDevice, s));
    p += pieces[i].bytes;
  }
  RTM_CUDA(cudaStreamSynchronize(s));
  return RTM_OK;
}
