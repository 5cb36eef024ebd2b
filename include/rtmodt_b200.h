/*
 * rtmodt_b200.h - C ABI of librtmodt_b200.so: the B200 (sm_100a) implementation of the
 * per-frame, post-backbone hot path of RTMODT, batched over many independent video streams.
 *
 * The reference (100 % Python, no FFI of its own) exposes this path as three classes called
 * once per frame from tools/run_pipeline.py:133,138,145.  Each entry point below replaces the
 * arithmetic behind one of those calls; the Python facades in
 * real-time-multi-object-detection---tracking-system_b200/ (Detector, MultiObjectTracker,
 * ZoneEventEngine - same names, arguments and errors as the reference) bind them with ctypes.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named in a struct or argument list is a DEVICE pointer unless the name
 *     starts with host_; the caller owns every buffer; nothing is allocated, freed or
 *     synchronised on the per-frame path (the first call that sees a workspace sets it up); all kernels are enqueued on `stream`
 *     (a cudaStream_t passed as void*, e.g. torch.cuda.current_stream().cuda_stream);
 *   - return value 0 = ok, negative = error (rtm_last_error() gives the message);
 *   - conditions only the device can detect (a table or list that would overflow its
 *     capacity) are reported in a per-stream `status` word, never silently truncated;
 *   - "stream b" always means video stream b of the batch, not a CUDA stream.
 */
#ifndef RTMODT_B200_H
#define RTMODT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTM_VERSION 100 /* 0.1.0 */

typedef void* rtm_cuda_stream; /* cudaStream_t */

enum {
  RTM_OK = 0,
  RTM_ERR_INVALID = -1, /* bad argument (null pointer, size out of range, ...) */
  RTM_ERR_CUDA = -2,    /* a CUDA runtime call or launch failed                 */
  RTM_ERR_UNSUPPORTED = -3
};

/* element types of tensors handed over by the host framework */
enum { RTM_F32 = 0, RTM_F16 = 1, RTM_BF16 = 2 };

/* bits of the per-stream status words */
enum {
  RTM_STATUS_TRACK_OVERFLOW = 1, /* live tracks would exceed rtm_track_table.capacity  */
  RTM_STATUS_DET_OVERFLOW = 2,   /* det_count[b] > det_stride                          */
  RTM_STATUS_CAND_OVERFLOW = 4,  /* NMS candidates exceed the workspace capacity       */
  RTM_STATUS_EVENT_OVERFLOW = 8, /* events of one step exceed event_stride             */
  RTM_STATUS_ZONE_LIMIT = 16,    /* a stream has more than 64 zones                    */
  RTM_STATUS_ASSIGN_LIMIT = 32   /* RTM_ASSIGN_OPTIMAL without (enough) assign_scratch: more than
                                    4096 admissible pairs in a stage, or a conflict component with
                                    more than 32 rows or columns; with scratch: more admissible
                                    pairs in a stage than the scratch was sized for        */
};

/* per-detection outcome of one tracker step (rtm_track_step: det_kind) */
enum { RTM_DET_NONE = 0, RTM_DET_STAGE1 = 1, RTM_DET_STAGE2 = 2, RTM_DET_BIRTH = 3 };

int rtm_version(void);
const char* rtm_last_error(void);
/* SM count and compute capability of the current device (fails unless it is sm_100). */
int rtm_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------------------------------
 * P1  letterbox + normalise.  Replaces ultralytics LetterBox + BasePredictor.preprocess as
 * reached from src/detection/detector.py:100-111: scale-preserving cv2.resize(INTER_LINEAR,
 * 8-bit fixed point) -> centred constant-114 border -> BGR->RGB -> HWC->CHW -> /255.
 *   frames   u8, B frames of src_h x src_w x 3 (BGR), row_stride / frame_stride in bytes
 *   out      (B, 3, out_h, out_w) of out_dtype (RTM_BF16 / RTM_F16 / RTM_F32), contiguous
 * The resize geometry (r = min(out_h/src_h, out_w/src_w), new size, padding) is derived
 * exactly as LetterBox does (auto=False, scaleup=True, center=True).
 * ---------------------------------------------------------------------------------------- */
int rtm_letterbox(const uint8_t* frames, int32_t num_streams, int32_t src_h, int32_t src_w,
                  int64_t row_stride, int64_t frame_stride, void* out, int32_t out_dtype,
                  int32_t out_h, int32_t out_w, rtm_cuda_stream stream);

/* The same with the resize geometry given by the caller: the source is resized to new_w x new_h
 * and placed at (left, top) of the out_w x out_h output, the rest is 114.  This is how the other
 * LetterBox variants are expressed - auto=True (the minimum stride-32 rectangle ultralytics uses
 * for .pt models: 1080p -> 384 x 640, 5040 anchors), scaleup=False, center=False. */
int rtm_letterbox_ex(const uint8_t* frames, int32_t num_streams, int32_t src_h, int32_t src_w,
                     int64_t row_stride, int64_t frame_stride, void* out, int32_t out_dtype,
                     int32_t out_h, int32_t out_w, int32_t new_h, int32_t new_w, int32_t top,
                     int32_t left, rtm_cuda_stream stream);

/* ------------------------------------------------------------------------------------------
 * D1 + N1..N3  head decode, candidate filter, class-aware NMS, rescale.  Replaces
 * ultralytics Detect._inference / DFL / dist2bbox, ops.non_max_suppression (which calls
 * torchvision.ops.nms) and ops.scale_boxes, all reached from detector.py:100-111; the
 * outputs are what Detector._parse (detector.py:117-129) copies to the host.
 * ---------------------------------------------------------------------------------------- */
typedef struct rtm_nms_params {
  double iou_thres;       /* suppress when (double)IoU_f32 > iou_thres, as torchvision.ops.nms
                             compares (default.yaml:37, 0.45)                                    */
  float conf_thres;       /* keep class prob > (float)conf: torch compares `f32 tensor > python
                             float` in float32 (default.yaml:36, 0.35)                           */
  int32_t max_det;        /* first max_det survivors              (default.yaml:38, 100)  */
  int32_t agnostic;       /* 0: boxes offset by class*7680 before NMS (default.yaml:40)   */
  int32_t num_classes;    /* nc, <= 256 (80)                                              */
  uint32_t class_mask[8]; /* bit c set = class c wanted (`classes=` argument; all ones = None) */
} rtm_nms_params;

/* bytes of scratch rtm_decode_nms / rtm_nms_pred / rtm_post_backbone_step need for num_streams
 * streams of num_anchors.  The workspace holds a ring of candidate lists (consecutive calls with the
 * same workspace rotate through it, which is what allows the head scan of one step to overlap the
 * post stage of the step before) and a small header of counters.  Give every stream batch its own
 * workspace, 256-byte aligned.  The library clears the header the first time it sees a workspace
 * address on a device (synchronously: not on the per-frame path) and keeps a little host-side
 * bookkeeping for it (a CUDA stream and events of its own once rtm_step_io.scan_async is used);
 * rtm_workspace_release drops that bookkeeping - call it before the memory is freed or reused. */
size_t rtm_nms_workspace_bytes(int32_t num_streams, int32_t num_anchors);
int rtm_workspace_release(void* workspace);

/*
 * head[l]  (B, 64 + nc, img_h/stride_l, img_w/stride_l), strides 8/16/32, contiguous NCHW,
 *          channels = 4 sides x 16 DFL bins (side-major l,t,r,b) then nc class logits
 * scale    (B, 5) f32: gain, pad_x, pad_y, src_w, src_h of ops.scale_boxes per stream, or
 *          NULL to leave boxes in letterbox pixels
 * outputs  per stream b, rows [0, det_count[b]) of the det_stride-row slabs, score order:
 *          det_xyxy (B, det_stride, 4) f32, det_conf f32, det_cls i32,
 *          det_anchor i32 (anchor index 0..A-1), det_keep i32 (the index torchvision.ops.nms
 *          returns: position in the conf/class-filtered candidate list); det_anchor / det_keep
 *          may be NULL.  det_stride >= max_det.
 * status   (B) i32, OR-ed with RTM_STATUS_* bits (caller zeroes it)
 */
int rtm_decode_nms(const void* head_p3, const void* head_p4, const void* head_p5,
                   int32_t head_dtype, int32_t num_streams, int32_t img_h, int32_t img_w,
                   const rtm_nms_params* params, const float* scale, float* det_xyxy,
                   float* det_conf, int32_t* det_cls, int32_t* det_anchor, int32_t* det_keep,
                   int32_t* det_count, int32_t det_stride, int32_t* status, void* workspace,
                   size_t workspace_bytes, rtm_cuda_stream stream);

/* Same post-process on an already decoded prediction tensor, the exact input of
 * ops.non_max_suppression: pred (B, 4 + nc, A) f32 = xywh (letterbox px) + class probs. */
int rtm_nms_pred(const float* pred, int32_t num_streams, int32_t num_anchors,
                 const rtm_nms_params* params, const float* scale, float* det_xyxy,
                 float* det_conf, int32_t* det_cls, int32_t* det_anchor, int32_t* det_keep,
                 int32_t* det_count, int32_t det_stride, int32_t* status, void* workspace,
                 size_t workspace_bytes, rtm_cuda_stream stream);

/* Head decode only (Detect._inference): pred (B, 4 + nc, A) f32, for tolerance checks. */
int rtm_decode_head(const void* head_p3, const void* head_p4, const void* head_p5,
                    int32_t head_dtype, int32_t num_streams, int32_t img_h, int32_t img_w,
                    int32_t num_classes, float* pred, rtm_cuda_stream stream);

/* ------------------------------------------------------------------------------------------
 * T1..T4  tracker step.  Replaces _ByteTrackCore.update / _batch_iou / _linear_assignment
 * (greedy branch) / _age_tracks, src/tracking/tracker.py:58-194.
 * One table = the `_tracks` lists + `_next_id` of B independent trackers, rows in creation
 * order (= ascending track_id), SoA, `capacity` rows per stream.
 * ---------------------------------------------------------------------------------------- */
typedef struct rtm_track_table {
  int32_t num_streams;
  int32_t capacity;
  int32_t* count;             /* (B)            live rows                      */
  int32_t* next_id;           /* (B)            tracker.py:55, starts at 1     */
  int32_t* track_id;          /* (B, capacity)                                  */
  float* xyxy;                /* (B, capacity, 4)                               */
  float* confidence;          /* (B, capacity)                                  */
  int32_t* class_id;          /* (B, capacity)                                  */
  int32_t* age;               /* (B, capacity)  number of matches + 1           */
  int32_t* time_since_update; /* (B, capacity)  1 = matched / born this step    */
} rtm_track_table;

/*
 * table_in -> table_out (two distinct tables; the caller ping-pongs them).
 * det_*        (B, det_stride[, 4]) detections of this frame, det_count (B)
 * det_track_id (B, det_stride) out, may be NULL: id of the track each detection updated or
 *              created, 0 = discarded; det_kind likewise with RTM_DET_*
 * src_row      (B, capacity) out, may be NULL: for every row of table_out the row of table_in
 *              it came from, -1 for a track born in this step
 * A stream with det_count == 0 only ages its tracks (tracker.py:70-73: no pruning).
 */
int rtm_track_step(const rtm_track_table* table_in, const rtm_track_table* table_out,
                   const float* det_xyxy, const float* det_conf, const int32_t* det_cls,
                   const int32_t* det_count, int32_t det_stride, float track_thresh,
                   float match_thresh, int32_t track_buffer, int32_t* det_track_id,
                   int32_t* det_kind, int32_t* src_row, int32_t* status,
                   rtm_cuda_stream stream);

/*
 * Row K of the scope table: opt-in motion model.  THE REFERENCE HAS NO KALMAN FILTER (its
 * tracker matches detections against the last matched box, tracker.py:93,112); with
 * kalman_in / kalman_out == NULL rtm_track_step_ex is rtm_track_step.  With them, every track
 * carries the constant-velocity filter of ByteTrack (ifzhang/ByteTrack, yolox/tracker/
 * kalman_filter.py: state x, y, a = w/h, h + velocities; std_weight_position 1/20,
 * std_weight_velocity 1/160) and the association of both stages runs against the box PREDICTED
 * for this frame instead of the stored one.  Everything else (thresholds, greedy assignment,
 * births, ageing, pruning, the stored xyxy = last matched detection) is unchanged.  With
 * diagonal initial covariance and diagonal process / measurement noise the 8 x 8 filter is four
 * independent (position, velocity) filters; that is how the state is stored:
 *   mean (B, capacity, 8)   x, y, a, h, vx, vy, va, vh
 *   cov  (B, capacity, 12)  per coordinate: var(pos), cov(pos, vel), var(vel)
 * float32 throughout; a track that was not matched in the previous step has vh zeroed before the
 * prediction (STrack.predict).
 */
typedef struct rtm_kalman_state {
  float* mean;
  float* cov;
} rtm_kalman_state;

enum { RTM_ASSIGN_GREEDY = 0, /* tracker.py:182-194, what the reference runs without `lap` */
       RTM_ASSIGN_OPTIMAL = 1 /* tracker.py:168-181: lap.lapjv(1 - IoU, extend_cost, cost_limit) */ };

typedef struct rtm_track_options {
  float track_thresh, match_thresh;
  int32_t track_buffer;
  int32_t assignment;                /* RTM_ASSIGN_* */
  const rtm_kalman_state* kalman_in; /* both NULL: no motion model (the reference) */
  const rtm_kalman_state* kalman_out;
  double cost_limit;                 /* RTM_ASSIGN_OPTIMAL: lap's cost_limit, `1 - match_thresh` evaluated in
                                        double as tracker.py:170 does; a pair is admissible iff
                                        (double)float32(1 - IoU) < cost_limit                         */
  /* RTM_ASSIGN_OPTIMAL: device scratch for association stages that outgrow the shared-memory solver (crowds, low
   * thresholds) - rtm_assign_scratch_bytes() bytes for the whole batch, or NULL / 0: such stages then set
   * RTM_STATUS_ASSIGN_LIMIT.  lap.lapjv itself has no such limits (tracker.py:168-181). */
  void* assign_scratch;
  size_t assign_scratch_bytes;
} rtm_track_options;

/* scratch for num_streams streams whose stages hold up to max_pairs admissible (track, detection) pairs each */
size_t rtm_assign_scratch_bytes(int32_t num_streams, int32_t capacity, int32_t det_stride, int32_t max_pairs);

int rtm_track_step_ex(const rtm_track_table* table_in, const rtm_track_table* table_out,
                      const float* det_xyxy, const float* det_conf, const int32_t* det_cls,
                      const int32_t* det_count, int32_t det_stride, const rtm_track_options* options,
                      int32_t* det_track_id, int32_t* det_kind, int32_t* src_row, int32_t* status,
                      rtm_cuda_stream stream);

/* ------------------------------------------------------------------------------------------
 * Z1..Z3  zone step.  Replaces ZoneEventEngine.process, src/events/zone_engine.py:82-132
 * (centroid, cv2.pointPolygonTest >= 0, dwell timer, cooldown ledger, stale purge).
 * ---------------------------------------------------------------------------------------- */
typedef struct rtm_zone_set {
  int32_t num_streams;
  int32_t num_columns;         /* state columns per track row (>= distinct zone names of any stream) */
  const int32_t* zone_offsets; /* (B + 1)  zones of stream b are [zone_offsets[b], zone_offsets[b+1]) */
  const int32_t* poly_offsets; /* (Z + 1)  vertices of zone z are [poly_offsets[z], poly_offsets[z+1]) */
  const int32_t* poly_xy;      /* (V, 2)   int32 vertices, zone_engine.py:143                          */
  const double* dwell_sec;     /* (Z)      zone_engine.py:148                                           */
  const double* cooldown_sec;  /* (Z)      zone_engine.py:149                                           */
  const int32_t* column;       /* (Z)      state column of zone z: zones of one stream that share a
                                           name share a column (the reference keys state by name)    */
} rtm_zone_set;

/* dwell / cooldown state attached to track rows: (B, num_columns, capacity) f64 each.
 * first_seen: NaN = not inside (zone_engine.py:96-99); last_alert: 0.0 = never (:105). */
typedef struct rtm_zone_state {
  double* first_seen;
  double* last_alert;
} rtm_zone_state;

typedef struct rtm_zone_event { /* 64 bytes; the deterministic fields of ZoneEvent (:29-45) */
  int32_t stream;
  int32_t frame_id;
  int32_t track_id;
  int32_t zone;     /* index of the zone within its stream */
  int32_t class_id;
  int32_t cx, cy;   /* int-truncated centroid, zone_engine.py:90-91 */
  int32_t row;      /* row of the track in `tracks` */
  double dwell;     /* now - first_seen (unrounded; the facade applies round(dwell, 2)) */
  double now;
  float xyxy[4];
} rtm_zone_event;

/*
 * tracks      table whose rows with time_since_update == 1 are the tracks passed to process();
 *             every other live row is "not present in this call": its dwell timers are purged
 *             (zone_engine.py:128-130), its cooldown entries are kept.
 * src_row     (B, capacity) row of state_in holding each row's previous state (-1 = none), or
 *             NULL for the identity; state_in / state_out may be the same object only then.
 * now         scalar clock (zone_engine.py:84) used when now_per_stream is NULL
 * events      (B, event_stride) records in (row, zone) order, event_count (B)
 */
int rtm_zone_step(const rtm_zone_set* zones, const rtm_track_table* tracks, const int32_t* src_row,
                  const rtm_zone_state* state_in, const rtm_zone_state* state_out, double now,
                  const double* now_per_stream, int32_t frame_id, rtm_zone_event* events,
                  int32_t event_stride, int32_t* event_count, int32_t* status,
                  rtm_cuda_stream stream);

/* ------------------------------------------------------------------------------------------
 * The whole post-backbone step for B streams: decode + NMS -> tracker -> zones, enqueued
 * back to back on `stream` (what run_pipeline.py:133-145 does per frame, minus the conv
 * forward).  Arguments as in the three calls above; det_* buffers carry the detections from
 * the first stage to the second and stay readable afterwards.
 * ---------------------------------------------------------------------------------------- */
typedef struct rtm_step_io {
  /* detector post-process */
  const void* head_p3;
  const void* head_p4;
  const void* head_p5;
  int32_t head_dtype, img_h, img_w;
  const float* scale;
  float* det_xyxy;
  float* det_conf;
  int32_t* det_cls;
  int32_t* det_anchor;
  int32_t* det_keep;
  int32_t* det_count;
  int32_t det_stride;
  void* workspace;
  size_t workspace_bytes;
  /* tracker */
  const rtm_track_table* table_in;
  const rtm_track_table* table_out;
  float track_thresh, match_thresh;
  int32_t track_buffer;
  int32_t* det_track_id;
  int32_t* det_kind;
  int32_t* src_row;
  /* zones */
  const rtm_zone_set* zones;
  const rtm_zone_state* state_in;
  const rtm_zone_state* state_out;
  double now;
  const double* now_per_stream;
  int32_t frame_id;
  rtm_zone_event* events;
  int32_t event_stride;
  int32_t* event_count;
  int32_t* status;
  /* opt-in motion model of the tracker stage (see rtm_track_step_ex); NULL = the reference */
  const rtm_kalman_state* kalman_in;
  const rtm_kalman_state* kalman_out;
  /* tracker assignment rule, RTM_ASSIGN_* (0 = greedy, what the reference runs without `lap`), and
   * lap's cost_limit for RTM_ASSIGN_OPTIMAL (see rtm_track_options) */
  int32_t assignment;
  double cost_limit;
  void* assign_scratch;        /* as in rtm_track_options */
  size_t assign_scratch_bytes;
  /* Pipelining of consecutive steps.  scan_async = 0 (default): everything goes to `stream` as ordinary
   * launches; the head tensors are ordered on `stream` like any other input and a step starts when the
   * step before it has finished.  scan_async = 1: the caller states that the head tensors are complete
   * once heads_ready_event (a cudaEvent_t, or NULL = complete already) has fired - e.g. the backbone runs
   * on its own stream and records an event per frame.  The step's kernel is then enqueued on a stream the
   * library owns, behind that event only, as a programmatic dependent of the previous step's kernel: its
   * head scan runs while the previous step's NMS / tracker / zone stage is still at work.  What must not be
   * overtaken is ordered on the device: a stream's post stage follows the same stream's previous one, and it
   * starts only once `stream` has reached this call (the result buffers and tables it overwrites may still be
   * read by work enqueued on `stream` before).  `stream` is made to wait for the kernel, so results are
   * ordered on `stream` exactly as in the default mode, and work enqueued on `stream` after the call is
   * ordered after the scan has read the head tensors.  (Do not enqueue cooperative launches that need the
   * whole GPU on `stream` between such calls: up to 64 post CTAs may be resident, waiting for the mark.) */
  int32_t scan_async;
  void* heads_ready_event;
  /* scan_async only.  0: the post stage of this step starts once `stream` has reached THIS call - it overwrites
   * result buffers (det_*, events, event_count) that work enqueued on `stream` since the previous call may still
   * read; as `stream` also waits for every step's kernel, the post stages of consecutive steps then follow each
   * other at kernel granularity.  1: the caller alternates between two sets of result buffers from call to call
   * (what this step overwrites was last read before the previous call), so the post stage only waits until
   * `stream` has reached the PREVIOUS call - consecutive post stages then follow each other stream by stream.
   * Either way the track tables / zone state / Kalman state the caller passes must not be touched by other work
   * on `stream` between scan_async calls without a synchronisation. */
  int32_t results_alternate;
} rtm_step_io;

int rtm_post_backbone_step(const rtm_step_io* io, const rtm_nms_params* params,
                           rtm_cuda_stream stream);

/*
 * The same step fed from HOST memory: copies this frame's head tensors host -> device into
 * the (device) head_p3/p4/p5 buffers of `io`, runs the step, and copies the per-step results
 * (events + counts, detections + counts, status) device -> host, all on `stream`, without
 * synchronising (the caller synchronises the stream or an event before reading host_*).
 * host_* buffers should be page-locked for the copies to be asynchronous.
 */
typedef struct rtm_step_host_io {
  const void* host_head_p3;
  const void* host_head_p4;
  const void* host_head_p5;
  rtm_zone_event* host_events; /* (B, host_event_stride or event_stride) */
  int32_t* host_event_count;   /* (B) */
  float* host_det_xyxy;        /* (B, det_stride, 4) or NULL */
  float* host_det_conf;        /* (B, det_stride)    or NULL */
  int32_t* host_det_cls;       /* (B, det_stride)    or NULL */
  int32_t* host_det_track_id;  /* (B, det_stride)    or NULL */
  int32_t* host_det_count;     /* (B)                or NULL */
  int32_t* host_status;        /* (B) */
  /* optional cross-stream ordering for double-buffered callers (cudaEvent_t as void*, or NULL):
   * `stream` waits on wait_event AFTER the host->device copies and BEFORE the kernels (the
   * previous step, running on another CUDA stream, owns the track / zone tables until then);
   * done_event is recorded after the last device->host copy. */
  void* wait_event;
  void* done_event;
  /* records per stream of host_events: only the first host_event_stride records of every stream's slab are
   * copied back (0 = all io->event_stride of them).  host_event_count tells how many a stream emitted; a count
   * beyond host_event_stride means the rest stayed on the device (io->events) for this step. */
  int32_t host_event_stride;
  /* optional ordering of the host->device copies themselves (cudaEvent_t as void*, or NULL): `stream` waits on
   * copy_wait_event BEFORE its copies and records copy_done_event right AFTER them.  A double-buffered caller passes
   * the previous step's copy_done_event as this step's copy_wait_event: the copies of consecutive steps then follow each
   * other on the link instead of running side by side - side by side they also end side by side, and the link idles
   * while both streams run their kernels (2 - 3 % of a PCIe-bound step). */
  void* copy_wait_event;
  void* copy_done_event;
} rtm_step_host_io;

int rtm_post_backbone_step_host(const rtm_step_io* io, const rtm_step_host_io* host_io,
                                const rtm_nms_params* params, rtm_cuda_stream stream);

/* ------------------------------------------------------------------------------------------
 * Checkpoint / resume of a batch's state (the reference keeps the same state in the Python
 * lists and dicts of _ByteTrackCore and ZoneEventEngine and has no way to save it): the track table with its
 * counters, the zone dwell / cooldown state and the Kalman state of all B streams, packed into ONE host blob
 * (a 64-byte header that records the shapes, then the arrays).  zone_state / kalman may be NULL (not stored;
 * num_columns is then ignored).  Both calls enqueue their copies on `stream` and synchronise it before they
 * return - they are not on the per-frame path.  rtm_state_import refuses a blob whose header does not match
 * the tables it is given.
 * ---------------------------------------------------------------------------------------- */
size_t rtm_state_bytes(int32_t num_streams, int32_t capacity, int32_t num_columns, int32_t with_zone_state,
                       int32_t with_kalman);
int rtm_state_export(const rtm_track_table* table, const rtm_zone_state* zone_state, int32_t num_columns,
                     const rtm_kalman_state* kalman, void* host_blob, size_t blob_bytes, rtm_cuda_stream stream);
int rtm_state_import(const rtm_track_table* table, const rtm_zone_state* zone_state, int32_t num_columns,
                     const rtm_kalman_state* kalman, const void* host_blob, size_t blob_bytes,
                     rtm_cuda_stream stream);

/* ------------------------------------------------------------------------------------------
 * Measurement aid (off by default; nothing below runs on the per-frame path unless enabled).
 * While enabled, every kernel launch of the library is bracketed by a pair of CUDA events on
 * the launching stream; rtm_profile_read synchronises them, adds the elapsed milliseconds and
 * launch counts per kernel kind into ms_sum[RTM_K_COUNT] / launches[RTM_K_COUNT] and clears
 * the record.  bench.py uses it for the per-kernel roofline figures.
 * ---------------------------------------------------------------------------------------- */
enum {
  RTM_K_LETTERBOX = 0,
  RTM_K_DECODE = 1, /* decode_candidates_kernel: the HBM-bound head scan */
  RTM_K_NMS = 2,
  RTM_K_TRACK = 3,
  RTM_K_ZONE = 4,
  RTM_K_PRED = 5,
  RTM_K_POST = 6, /* fused NMS + tracker + zones of rtm_post_backbone_step (two-launch path) */
  RTM_K_STEP = 7, /* step_kernel: head scan + NMS + tracker + zones in one launch */
  RTM_K_COUNT = 8
};
int rtm_profile_enable(int32_t on);
int rtm_profile_read(double* ms_sum, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* RTMODT_B200_H */
