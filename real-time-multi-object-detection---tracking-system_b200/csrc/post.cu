// rtm_post_backbone_step: the whole post-backbone step of B streams.
//
// One launch (step_kernel, the default wherever the tiled head scan applies): a grid of scan CTAs followed by
// post CTAs, one CTA per SM.
//   scan CTAs   the TMA head scan (decode_body.cuh): three teams of consumer warps behind one producer lane,
//               tiles handed out by tickets                                                         HBM-bound
//   post CTAs   NMS -> tracker step -> zone step of one stream at a time (the same stage bodies the stand-alone
//               kernels run: nms_body.cuh, track_body.cuh, zone_body.cuh), started once the launch's scan CTAs
//               have reported completion through a counter in the workspace header
// Post CTAs come last in the grid, so they move in as the scan CTAs retire - no second launch, no launch latency
// between the stages.  Consecutive steps overlap when the caller has declared the head tensors complete
// (rtm_step_io.scan_async): the kernel then runs on a stream of the library's own as a programmatic dependent of
// the previous step's kernel, every CTA releases the dependent at once, and what the next step's CTAs may not
// overtake is ordered on the device instead - a per-stream sequence number (a stream's post stage waits for the
// same stream's previous one), a per-slot counter (a scan may refill a slot of the candidate ring only when its
// last readers are done) and a mark the caller's stream writes when it reaches the call (the post stage
// overwrites result buffers the caller may still be reading).  The post CTAs of step k are resident before any
// CTA of step k+1 starts, so the serial stages never wait for SMs; the next scan fills the SMs they leave.
//
// Two launches (fallback: shapes the tiled scan does not cover, RTM_STEP_FUSED=0, diagnosis builds):
//   1. decode_tma / decode_scan (nms.cu)  head scan, candidate lists
//   2. post_kernel (this file)            one CTA per stream: NMS -> tracker step -> zone step
#include <stdlib.h>

#include <chrono>
#include <stdio.h>
#include "decode_body.cuh"
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

// NMS -> tracker -> zones of stream b by the whole CTA; ends with every thread's global writes issued.
// Only the tracker and zone stages depend on the stream's previous step (its track table and zone state): `chain`
// is called between the NMS and them - the step kernel waits there for that step to have left its tables, so
// that the NMS of a step runs beside the tracker / zone stages of the step before.  What the later stages need
// that does not depend on the previous step (zone table, polygons) is fetched first, behind the NMS; the head of
// the track table right after `chain` (without one: behind the NMS as well).
struct NoChain {
  static constexpr bool kWaits = false;
  __device__ __forceinline__ void operator()() const {}
};

template <bool WITH_OPTIMAL, typename Chain = NoChain>
__device__ __forceinline__ void post_stream(const PostArgs& a, const int b, unsigned char* smem_raw, int* s_keep, int* s_scan,
                                            Chain chain = Chain()) {
  rtm::TrackPrefetch* tpf = reinterpret_cast<rtm::TrackPrefetch*>(smem_raw + a.work_bytes);
  rtm::ZonePrefetch* zpf = reinterpret_cast<rtm::ZonePrefetch*>(tpf + 1);
  RTM_TL(0);
  auto prefetch = [&]() {
    if (a.has_zones) rtm::zone_prefetch<kPostThreads>(a.zone, b, zpf);
    if (!Chain::kWaits) rtm::track_prefetch<kPostThreads>(a.trk, b, tpf);
  };
  rtm::nms_stream<kPostThreads>(a.ws, a.prm, a.iou_gate, a.out, b, smem_raw, s_keep, s_scan, prefetch);  // ends with a barrier
  RTM_TL(10);
  if (Chain::kWaits) {
    chain();  // contains a block barrier
    rtm::track_prefetch<kPostThreads>(a.trk, b, tpf);
    __syncthreads();
  }
  rtm::track_stream<kPostThreads, WITH_OPTIMAL>(a.trk, b, smem_raw, tpf);
  __syncthreads();
  RTM_TL(20);
  if (a.has_zones) rtm::zone_stream<kPostThreads>(a.zone, b, smem_raw, s_scan, zpf);
  RTM_TL(30);
}

template <bool WITH_OPTIMAL>
__global__ void __launch_bounds__(kPostThreads, RTM_POST_CTAS_PER_SM) post_kernel(const __grid_constant__ PostArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_keep[rtm::kMaxDetCap];
  __shared__ int s_scan[33];
  post_stream<WITH_OPTIMAL>(a, blockIdx.x, smem_raw, s_keep, s_scan);
}

// ---------------------------------------------------------------------------------------
// The one-launch step
// ---------------------------------------------------------------------------------------
// Two step CTAs share an SM (at most half its shared memory and 64 registers per thread each), so that an SM whose
// one CTA walks through the latency-bound post stages keeps scanning with the other.  Every CTA scans first - two
// teams of five consumer warps, each behind its own producer warp (12 warps; the block's other four wait at the
// barrier that ends the scan) - and when the tickets run out it turns to the post stage: it draws stream numbers
// from a second ticket counter until that runs out too.  So the grid is exactly the GPU's CTA slots, every CTA of
// a launch is resident from the start, and the next launch (a programmatic dependent) moves in CTA by CTA as this
// one's scanners retire, while the few CTAs that drew a stream finish their post stages beside it.
constexpr int kScanTeams = 2;
constexpr int kStepThreads = kPostThreads;
constexpr int kStepCtasPerSm = 2;
constexpr int kScanRoleThreads = rtm::tma_threads(rtm::kStepTileW, kScanTeams);
static_assert(kScanRoleThreads <= kStepThreads, "the scan role must fit the step kernel's block");

struct StepArgs {
  PostArgs post;  // post.ws describes the candidate slot; its tile_counter is null (the post stage re-arms nothing)
  rtm::TmaGeom tg;
  float logit_gate;
  int num_streams;
  int post_workers;  // CTAs 0 .. post_workers - 1 turn to the post stage after their scan; the others leave
  int seq;           // step kernels launched on this workspace before this one
  // device-side ordering (see the header of this file); all counters live in the workspace header, all are running
  // totals, and the host knows the value each will have reached when the thing waited for has happened
  int* tile_tickets;       // this launch's tile-ticket counter (one of a ring of 64, zero when the launch starts)
  int* rearm;              // the counter 32 launches ahead: zeroed by this launch
  int* tiles_done;         // slot counter: tiles of the slot's scans finished
  int tiles_done_target;   //   its value once this launch's scan is complete
  int* slot_free;          // slot counter: streams whose post stage is done with the slot
  int slot_free_target;    //   its value once the slot's PREVIOUS readers were done (the scan waits for it)
  int* post_ticket;        // slot counter: post-stage tickets drawn; ticket - post_ticket_base = stream, until >= num_streams
  int post_ticket_base;
  const int* caller_mark;  // written by the caller's stream as it reaches a call; 0 target = not waited for
  int caller_target;
  int* stream_seq;         // (B) steps completed per stream
  int chain;               // a stream's post stage waits for stream_seq[b] == seq (the previous step may still run)
  const void* head[3];     // the head tensors (LAZY: the decoder warps read the DFL values of candidates from them)
};

// Diagnosis builds (-DRTM_TIMELINE, tools/step_timeline.py): CTA-level stamps of the step kernel, a row of 8 words per
// (launch % 16, CTA): [0] start, [1] scan over, [2] end, [3] streams post-processed, [4] SM, [5] first post stage starts
#ifdef RTM_TIMELINE
__device__ __forceinline__ void step_mark(const int seq, const int i, const unsigned long long v = ~0ull) {
  if (threadIdx.x == 0 && rtm::g_timeline) {
    unsigned long long t = v;
    if (v == ~0ull) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    rtm::g_timeline[16384 + (static_cast<size_t>(seq & 15) * 512 + blockIdx.x) * 8 + i] = t;  // behind the per-stage rows
  }
}
#define RTM_STEP_MARK(seq, ...) step_mark(seq, __VA_ARGS__)
#else
#define RTM_STEP_MARK(seq, ...) ((void)0)
#endif

// the tracker / zone stages of a stream follow the same stream's previous step (it may still be running)
struct StreamChain {
  const int* seq_word;
  int target, on;
  static constexpr bool kWaits = true;
  __device__ __forceinline__ void operator()() const {
    if (threadIdx.x == 0 && on) rtm::spin_until_ge(seq_word, target);
    __syncthreads();
  }
};

template <typename T, bool WITH_OPTIMAL, bool LAZY>
__global__ void __launch_bounds__(kStepThreads, kStepCtasPerSm) step_kernel(const __grid_constant__ rtm::TmaMaps maps,
                                                                            const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(16) rtm::ScanCtl ctl;
  __shared__ int s_keep[rtm::kMaxDetCap];
  __shared__ int s_scan[33];
  __shared__ int s_job;
  __shared__ __align__(16) unsigned char cq_raw[LAZY ? sizeof(rtm::CandQueue) : 16];
  rtm::CandQueue* const cq = reinterpret_cast<rtm::CandQueue*>(cq_raw);
  const int tid = threadIdx.x;
  // the next step's kernel (a programmatic dependent on the library's stream) may start as soon as every CTA of
  // this grid has got here: what it must not overtake is ordered through the counters above
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  RTM_STEP_MARK(a.seq, 0);
#ifdef RTM_TIMELINE
  {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    RTM_STEP_MARK(a.seq, 4, static_cast<unsigned long long>(smid));
  }
#endif
  if (blockIdx.x == 0 && tid == 0) *a.rearm = 0;

  // ---- scan ----
  if (LAZY) {
    rtm::cand_queue_init(cq);
    __syncthreads();
  }
  const rtm::ScanSync scan_sync{a.slot_free, a.slot_free_target, a.tiles_done};
  if (tid < kScanRoleThreads) {
    rtm::Workspace ws = a.post.ws;
    ws.tile_counter = a.tile_tickets;
    rtm::tma_scan_cta<T, true, rtm::kStepTileW, kScanTeams, LAZY>(maps, a.tg, a.post.prm, a.logit_gate, ws, scan_sync,
                                                     static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), smem_raw, &ctl, cq);
  } else if (LAZY) {
    // the block's other warps decode the candidates the scan finds (D1), until the scan is over and the queue empty
    rtm::cand_decoder_warp<T>(cq, a.head, a.tg, a.post.ws);
  }
  __syncthreads();  // the ring is shared memory for the post stage from here on
  if (LAZY && tid == 0) rtm::scan_publish(scan_sync, &ctl, kScanTeams);  // (after the decoders' stores)
  RTM_STEP_MARK(a.seq, 1);
  RTM_STEP_MARK(a.seq, 6, (ctl.tl[0] & 0xffffffffull) | (ctl.tl[1] << 32));  // team 0: consumer wait ns | consumer loop ns
  RTM_STEP_MARK(a.seq, 7, (ctl.tl[3] & 0xffffffffull) | (ctl.tl[2] << 32));  // team 0: producer wait ns | tiles

  // ---- post: streams by ticket, on the launch's first CTAs ----
  // (as many as there are streams, 64 at most: with more, the post stages of a large batch would take every CTA slot
  // at once when its scan ends and the next launch could not start scanning beside them)
  int done = 0;
  while (static_cast<int>(blockIdx.x) < a.post_workers) {
    if (tid == 0) s_job = atomicAdd(a.post_ticket, 1) - a.post_ticket_base;
    __syncthreads();
    const int b = s_job;
    if (b >= a.num_streams) break;
    if (tid == 0 && done == 0) {
      rtm::spin_until_ge(a.tiles_done, a.tiles_done_target);                    // this launch's candidate lists are complete
      if (a.caller_target) rtm::spin_until_ge(a.caller_mark, a.caller_target);  // the caller's stream is far enough
    }
    if (done == 0) RTM_STEP_MARK(a.seq, 5);
    __syncthreads();
    post_stream<WITH_OPTIMAL>(a.post, b, smem_raw, s_keep, s_scan, StreamChain{a.stream_seq + b, a.seq, a.chain});
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicExch(a.stream_seq + b, a.seq + 1);
      atomicAdd(a.slot_free, 1);
    }
    ++done;
  }
  RTM_STEP_MARK(a.seq, 2);
  RTM_STEP_MARK(a.seq, 3, static_cast<unsigned long long>(done));
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


#ifdef RTM_PROBES
// tools/probe_occupy.py: holds `ctas` SMs (one CTA of `smem` bytes each) for `ns` nanoseconds
__global__ void occupy_kernel(long long ns) {
  extern __shared__ unsigned char occ_smem[];
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (threadIdx.x == 0) occ_smem[0] = 1;
  do {
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (static_cast<long long>(t - t0) < ns && static_cast<long long>(t - t0) < 100000000ll);
}
}  // namespace
extern "C" int rtm_debug_occupy(int ctas, int smem, long long ns, void* stream) {
  RTM_CUDA(cudaFuncSetAttribute(occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  occupy_kernel<<<ctas, 128, smem, static_cast<cudaStream_t>(stream)>>>(ns);
  RTM_LAUNCH_CHECK("occupy_kernel");
  return RTM_OK;
}
namespace {
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

bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  return e && *e ? e[0] != '0' : dflt;
}

bool fuse_enabled() {
  static const bool v = env_flag("RTM_FUSE_POST", true);
  return v;
}

// RTM_STEP_FUSED=0: the two-launch pipeline even where the one-launch step applies
bool step_fused_enabled() {
  static const bool v = env_flag("RTM_STEP_FUSED", true);
  return v;
}

int env_int_or(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

// cuStreamWriteValue32 through the runtime's driver entry point table (no link dependency on libcuda)
typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWriteValue32Fn stream_write_value32() {
  static StreamWriteValue32Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StreamWriteValue32Fn>(p);
  }
  return fn;
}

template <typename T, bool LAZY>
int launch_step_kernel_as(const cudaLaunchConfig_t& cfg, bool optimal, const rtm::TmaMaps& maps, const StepArgs& a, size_t smem) {
  if (int rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(optimal ? step_kernel<T, true, LAZY> : step_kernel<T, false, LAZY>), smem))
    return rc;
  if (optimal) RTM_CUDA(cudaLaunchKernelEx(&cfg, step_kernel<T, true, LAZY>, maps, a));
  else RTM_CUDA(cudaLaunchKernelEx(&cfg, step_kernel<T, false, LAZY>, maps, a));
  return RTM_OK;
}
template <typename T>
int launch_step_kernel(const cudaLaunchConfig_t& cfg, bool optimal, bool lazy, const rtm::TmaMaps& maps, const StepArgs& a, size_t smem) {
  return lazy ? launch_step_kernel_as<T, true>(cfg, optimal, maps, a, smem) : launch_step_kernel_as<T, false>(cfg, optimal, maps, a, smem);
}
template <typename T>
const void* step_kernel_ptr(bool optimal, bool lazy) {
  if (lazy) return optimal ? reinterpret_cast<const void*>(step_kernel<T, true, true>) : reinterpret_cast<const void*>(step_kernel<T, false, true>);
  return optimal ? reinterpret_cast<const void*>(step_kernel<T, true, false>) : reinterpret_cast<const void*>(step_kernel<T, false, false>);
}
// static shared memory of a step kernel variant (the ring depth is chosen so that two CTAs share an SM)
size_t step_kernel_static_smem(int head_dtype, bool optimal, bool lazy) {
  static size_t cache[3][2][2] = {};
  const int d = head_dtype == RTM_F32 ? 0 : (head_dtype == RTM_F16 ? 1 : 2);
  size_t& v = cache[d][optimal][lazy];
  if (!v) {
    const void* f = d == 0 ? step_kernel_ptr<float>(optimal, lazy) : (d == 1 ? step_kernel_ptr<__half>(optimal, lazy) : step_kernel_ptr<__nv_bfloat16>(optimal, lazy));
    cudaFuncAttributes attr;
    if (cudaFuncGetAttributes(&attr, f) != cudaSuccess) {
      (void)cudaGetLastError();
      return 16 * 1024;
    }
    v = attr.sharedSizeBytes;
  }
  return v;
}

void fill_post_args(PostArgs* a, const rtm_step_io* io, const rtm_nms_params* params, size_t work) {
  a->prm = *params;
  a->iou_gate = rtm::iou_gate_for(params->iou_thres);
  a->out = rtm::NmsOut{io->scale, io->det_xyxy, io->det_conf, io->det_cls, io->det_anchor, io->det_keep, io->det_count,
                       io->det_stride, io->status};
  a->trk = rtm::TrackArgs{*io->table_in, *io->table_out, io->det_xyxy, io->det_conf, io->det_cls, io->det_count,
                          io->det_stride, io->track_thresh, io->match_thresh, io->track_buffer, io->det_track_id,
                          io->det_kind, io->src_row, io->status,
                          nullptr, nullptr, nullptr, nullptr, io->assignment, io->cost_limit, nullptr, 0};
  if (io->assign_scratch && io->assign_scratch_bytes) {
    a->trk.assign_scratch = static_cast<unsigned char*>(io->assign_scratch);
    a->trk.assign_scratch_per_stream = io->assign_scratch_bytes / io->table_in->num_streams / 16 * 16;
  }
  if (io->kalman_in) {
    a->trk.kf_mean_in = io->kalman_in->mean;
    a->trk.kf_cov_in = io->kalman_in->cov;
    a->trk.kf_mean_out = io->kalman_out->mean;
    a->trk.kf_cov_out = io->kalman_out->cov;
  }
  a->has_zones = io->zones != nullptr;
  if (a->has_zones)
    a->zone = rtm::ZoneArgs{*io->zones, *io->table_out, io->src_row, *io->state_in, *io->state_out, io->now,
                            io->now_per_stream, io->frame_id, io->events, io->event_stride, io->event_count,
                            io->status};
  a->work_bytes = static_cast<int>(work);
}

// RTM_HOST_TIMING=1: where the host side of a step goes (std::chrono between the sections of step_fused, printed every
// 1000 calls)
struct HostTiming {
  bool on = env_flag("RTM_HOST_TIMING", false);
  std::chrono::steady_clock::time_point t;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  long calls = 0;
  void start() {
    if (on) t = std::chrono::steady_clock::now();
  }
  void lap(int i) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    acc[i] += std::chrono::duration<double, std::micro>(n - t).count();
    t = n;
  }
  void end() {
    if (!on || ++calls % 1000) return;
    fprintf(stderr, "rtm host timing, us per step: plan %.2f  workspace %.2f  args %.2f  caller mark %.2f  launch %.2f  done event %.2f\n",
            acc[0] / 1000, acc[1] / 1000, acc[2] / 1000, acc[3] / 1000, acc[4] / 1000, acc[5] / 1000);
    for (double& a : acc) a = 0;
  }
};

// The one-launch step.  Returns 1 when it was enqueued, 0 when the caller should take the two-launch path.
int step_fused(const rtm_step_io* io, const rtm_nms_params* params, cudaStream_t s, size_t work, size_t post_smem) {
  const int B = io->table_in->num_streams;
  const bool optimal = io->assignment == RTM_ASSIGN_OPTIMAL;
  static HostTiming ht;
  ht.start();
  rtm::TmaScanPlan plan;
  int rc = rtm::plan_tma_scan80(io->head_p3, io->head_p4, io->head_p5, io->head_dtype, B, io->img_h, io->img_w, params, &plan);
  if (rc <= 0) return rc;
  ht.lap(0);
  if (!plan.nc80) return 0;
  // ring depth: as many stages as fit into a CTA's share of the SM (two CTAs per SM), the same number for every team
  static const int stages_env = env_int_or("RTM_STEP_STAGES", 0);
  static const bool lazy_env = env_flag("RTM_STEP_LAZY", true);
  const bool lazy = lazy_env && plan.tg.cls_tile_bytes % 128 == 0 && kStepThreads > kScanRoleThreads;
  // dynamic shared memory a CTA may use with a second CTA beside it: half the SM's 228 KB less 1 KB the system keeps per CTA
  // and the kernel's static part
  const size_t budget = (228 * 1024 / kStepCtasPerSm - 1024 - step_kernel_static_smem(io->head_dtype, optimal, lazy)) / 128 * 128;
  if (post_smem > budget) return 0;  // (large tables: the two-launch pipeline has the shared memory for them)
  // lazy box rows (RTM_STEP_LAZY=0 turns it off): the ring holds class rows only; candidates go through a queue to the
  // block's other warps, which read the 64 DFL values of each from global memory and decode them beside the scan
  const size_t stage_bytes = lazy ? plan.tg.cls_tile_bytes : plan.tg.tile_bytes;
  int stages = stages_env > 0 ? stages_env : static_cast<int>(budget / stage_bytes);
  if (stages > rtm::kMaxStages) stages = rtm::kMaxStages;
  // (lazy: eight stages fit, but with them the SM's whole shared memory is taken and the L1 that is left slows the post
  // stages: 28.1 us per step against 23.8 with six; four: 25.6)
  if (lazy && stages_env <= 0 && stages > 6) stages = 6;
  stages -= stages % kScanTeams;
  if (stages < kScanTeams) return 0;
  plan.tg.stages = stages;

  // first ring round static (no ticket round trip before the first load; every CTA of the grid starts at once, see the
  // grid size below), tickets after that (RTM_STEP_STATIC=0: tickets from the first tile on)
  static const int static_env = env_int_or("RTM_STEP_STATIC", 1);
  plan.tg.static_rounds = static_env;
  // RTM_STEP_L2_AHEAD: L2 prefetch distance in tiles (experiments; off by default: measured slower at every distance,
  // 34.9 - 38.1 vs 29.6 us per step - the prefetches double the TMA unit's row requests)
  static const int ahead_env = env_int_or("RTM_STEP_L2_AHEAD", 0);
  plan.tg.l2_ahead = ahead_env > 0 ? ahead_env : 0;
  const size_t ring = static_cast<size_t>(stages) * stage_bytes;
  const size_t smem = ring > post_smem ? ring : post_smem;

  rtm::WorkspaceCtx* ctx = nullptr;
  rc = rtm::workspace_ctx(io->workspace, io->workspace_bytes, B, s, &ctx);
  if (rc) return rc;
  StepArgs a;
  rc = rtm::take_scan_slot(ctx, io->workspace, io->workspace_bytes, B, plan.tg.g.num_anchors, &a.post.ws);
  if (rc) return rc;
  const int slot = a.post.ws.slot;
  ht.lap(1);
  fill_post_args(&a.post, io, params, work);
  a.tg = plan.tg;
  a.logit_gate = plan.logit_gate;
  a.num_streams = B;
  a.seq = ctx->seq++;
  // every CTA scans, the first `post_workers` of them then work through the post stages.  A launch that runs by
  // itself (ordinary launch) takes all CTA slots (59.2 us per step with the workers' slots left free, 58.4 with all).  In a chain of programmatic
  // dependent launches the grid is 3/8 of the slots: every CTA of a launch must have started before the next launch
  // can, so with small grids two or three launches are resident at any time and their scans overlap continuously -
  // no ramp and tail per launch (measured, 64 streams: 232 CTAs 29.2 us per step, 148: 28.3, 111: 27.6, 64: 26.9;
  // in a 20-step region 31.8 / 31.1 / 30.9 / 31.4)
  static const int grid_env = env_int_or("RTM_STEP_GRID", 0), workers_env = env_int_or("RTM_STEP_POST_CTAS", 64);
  a.post_workers = max(1, min(B, workers_env));
  const int slots = rtm::sm_count() * kStepCtasPerSm;
  const int want = grid_env > 0 ? grid_env : (io->scan_async ? max(slots * 3 / 8, a.post_workers) : slots);
  const int grid = max(a.post_workers, min((plan.tg.total_tiles + kScanTeams - 1) / kScanTeams, want));
  int* slot_words = a.post.ws.tile_counter;  // the slot's 32 header words
  a.post.ws.tile_counter = nullptr;
  a.tile_tickets = a.post.ws.sync + rtm::kSyncTicketRing + (a.seq & 63);
  a.rearm = a.post.ws.sync + rtm::kSyncTicketRing + ((a.seq + 32) & 63);
  a.tiles_done = slot_words + 1;
  a.slot_free = slot_words + 2;
  a.post_ticket = slot_words + 3;
  a.slot_free_target = ctx->slot_free_target[slot];
  ctx->slot_free_target[slot] += B;
  ctx->tiles_done_target[slot] += plan.tg.total_tiles;
  a.tiles_done_target = ctx->tiles_done_target[slot];
  a.post_ticket_base = ctx->post_ticket_base[slot];
  ctx->post_ticket_base[slot] += B + a.post_workers;  // every stream is drawn once, and every worker ends on a ticket past the last stream
  a.caller_mark = a.post.ws.sync;
  a.caller_target = 0;
  a.stream_seq = a.post.ws.sync + rtm::kSyncStreamSeq;
  a.head[0] = io->head_p3;
  a.head[1] = io->head_p4;
  a.head[2] = io->head_p5;

  const bool async = io->scan_async != 0;
  cudaStream_t ls = s;
  bool pdl = false;
  ht.lap(2);
  if (async) {
    rc = rtm::workspace_streams(ctx);
    if (rc) return rc;
    ls = ctx->stream;
    // (waits that are already satisfied are not enqueued: a wait between two kernels keeps the second from
    // being launched ahead as a programmatic dependent)
    if (io->heads_ready_event && cudaEventQuery(static_cast<cudaEvent_t>(io->heads_ready_event)) != cudaSuccess)
      RTM_CUDA(cudaStreamWaitEvent(ls, static_cast<cudaEvent_t>(io->heads_ready_event), 0));
    (void)cudaGetLastError();  // cudaEventQuery reports "not ready" through the error state
    // the post stage overwrites tables and result buffers the caller's stream may still be working with: it waits
    // (on the device) for a mark that stream writes when it gets here; the scan does not
    if (StreamWriteValue32Fn wv = stream_write_value32()) {
      const int mark = ++ctx->caller_seq;
      // results_alternate: the caller switches between two sets of result buffers from call to call, so what this
      // step overwrites was last read before the PREVIOUS call - the post stage need not wait for this call's mark,
      // which sits behind the caller stream's wait for the previous step's kernel
      a.caller_target = io->results_alternate ? mark - 1 : mark;
      const CUresult r = wv(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(a.post.ws.sync),
                            static_cast<cuuint32_t>(mark), 0);
      if (r != CUDA_SUCCESS) {
        rtm::set_error("cuStreamWriteValue32 failed (%d)", static_cast<int>(r));
        return RTM_ERR_CUDA;
      }
    } else {
      RTM_CUDA(cudaEventRecord(ctx->caller_mark, s));
      RTM_CUDA(cudaStreamWaitEvent(ls, ctx->caller_mark, 0));
    }
    // a programmatic dependent of the previous step kernel on the library's stream (same batch: the per-stream
    // sequence numbers line up); the first kernel of a sequence is an ordinary launch
    pdl = ctx->chain_streams == B && rtm::pdl_enabled() && !rtm::g_profile_on;
    ctx->chain_streams = B;
  } else if (ctx->stream && ctx->chain_streams) {
    // a step on the caller's stream after steps on the library's: those are complete as far as `s` is concerned
    // (it waited for each of them), and the next asynchronous step starts a new chain
    ctx->chain_streams = 0;
  }
  a.chain = pdl ? 1 : 0;
  ht.lap(3);

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kStepThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ls;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  {
    rtm::ProfileScope prof(RTM_K_STEP, ls);
    switch (io->head_dtype) {
      case RTM_F32:
        rc = launch_step_kernel<float>(cfg, optimal, lazy, plan.maps, a, smem);
        break;
      case RTM_F16:
        rc = launch_step_kernel<__half>(cfg, optimal, lazy, plan.maps, a, smem);
        break;
      default:
        rc = launch_step_kernel<__nv_bfloat16>(cfg, optimal, lazy, plan.maps, a, smem);
    }
  }
  if (rc) return rc;
  RTM_LAUNCH_CHECK("step_kernel");
  ht.lap(4);
  if (async) {
    RTM_CUDA(cudaEventRecord(ctx->done[slot], ls));
    RTM_CUDA(cudaStreamWaitEvent(s, ctx->done[slot], 0));
  }
  ht.lap(5);
  ht.end();
  return 1;
}

}  // namespace

extern "C" int rtm_post_backbone_step(const rtm_step_io* io, const rtm_nms_params* params,
                                      rtm_cuda_stream stream) {
  RTM_REQUIRE(io && params, "rtm_post_backbone_step: null argument");
  RTM_REQUIRE(io->table_in && io->table_out, "rtm_post_backbone_step: null track table");
  const int B = io->table_in->num_streams;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::recursive_mutex> lock(rtm::api_mutex());

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
                                io->kalman_out, io->cost_limit, io->assign_scratch, io->assign_scratch_bytes};
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

  if (step_fused_enabled()) {
    const int fused = step_fused(io, params, s, work, smem);
    if (fused != 0) return fused < 0 ? fused : RTM_OK;
  }

  // ---- two launches: head scan, then the post kernel ----
  PostArgs a;
  rtm::WorkspaceCtx* ctx = nullptr;
  int rc = rtm::workspace_ctx(io->workspace, io->workspace_bytes, B, s, &ctx);
  if (rc) return rc;
  ctx->chain_streams = 0;  // not a step kernel: the next one starts a new chain
  if (io->scan_async) {
    // the scan goes to the library's own stream, behind "heads ready" and behind the post kernel that last
    // read the slot it is about to fill; the caller's stream waits for it before the post kernel
    rc = rtm::workspace_streams(ctx);
    if (rc) return rc;
    const int slot = ctx->next_slot;
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
                                  io->workspace, io->workspace_bytes, &a.ws, s, nullptr, ctx->stream);
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
    ctx->covered = 0;  // this scan is not on the scan stream: the next asynchronous one orders itself afresh
  }
  fill_post_args(&a, io, params, work);
  rc = rtm::ensure_dynamic_smem(reinterpret_cast<const void*>(optimal ? post_kernel<true> : post_kernel<false>), smem);
  if (rc) return rc;
  {
    rtm::ProfileScope prof(RTM_K_POST, s);
#ifdef RTM_PROBES
    if (probe_bits() & 4)  // timing probe only: scans alone (the NMS stage would have re-armed the slot's ticket counter)
      RTM_CUDA(cudaMemsetAsync(a.ws.tile_counter, 0, sizeof(int), s));
    else
#endif
    if (optimal) post_kernel<true><<<B, kPostThreads, smem, s>>>(a);
    else post_kernel<false><<<B, kPostThreads, smem, s>>>(a);
  }
  RTM_LAUNCH_CHECK("post_kernel");
  if (io->scan_async) {
    RTM_CUDA(cudaEventRecord(ctx->consumed[a.ws.slot], s));
    ctx->consumed_valid[a.ws.slot] = true;
  }
  return RTM_OK;
}
