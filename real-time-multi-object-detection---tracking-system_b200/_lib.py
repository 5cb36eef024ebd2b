"""ctypes binding of librtmodt_b200.so (the C ABI declared in include/rtmodt_b200.h).

There is no CPU fallback: if the library is missing or no sm_100 device is present, every
compute entry point raises.  PyTorch is used only for device memory and CUDA streams.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
#: RTM_LIB_PATH selects another build of the same library (diagnosis builds of tools/post_timeline.py)
LIB_PATH = os.environ.get("RTM_LIB_PATH") or os.path.join(HERE, "librtmodt_b200.so")

RTM_F32, RTM_F16, RTM_BF16 = 0, 1, 2
STATUS_TRACK_OVERFLOW, STATUS_DET_OVERFLOW, STATUS_CAND_OVERFLOW = 1, 2, 4
STATUS_EVENT_OVERFLOW, STATUS_ZONE_LIMIT, STATUS_ASSIGN_LIMIT = 8, 16, 32
DET_NONE, DET_STAGE1, DET_STAGE2, DET_BIRTH = 0, 1, 2, 3
MAX_ZONES_PER_STREAM = 64
ZONE_STEP_EVENTS = 4096     # kZoneStepEvents: events one stream can emit in one step

_STATUS_TEXT = {
    STATUS_TRACK_OVERFLOW: "live tracks exceed the track table capacity (raise max_tracks)",
    STATUS_DET_OVERFLOW: "more detections than detection slots (raise max_dets)",
    STATUS_CAND_OVERFLOW: "NMS candidates exceed the workspace capacity",
    STATUS_EVENT_OVERFLOW: "zone events of one step exceed the event buffer (raise max_events)",
    STATUS_ZONE_LIMIT: f"a stream has more than {MAX_ZONES_PER_STREAM} zones",
    STATUS_ASSIGN_LIMIT: "optimal assignment: more admissible pairs in a stage than the solver's scratch holds "
                         "(raise max_pairs)",
}

i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


class RtmError(RuntimeError):
    """An rtm_* entry point returned an error code."""


class NmsParams(C.Structure):
    _fields_ = [("iou_thres", C.c_double), ("conf_thres", C.c_float), ("max_det", C.c_int32),
                ("agnostic", C.c_int32), ("num_classes", C.c_int32), ("class_mask", C.c_uint32 * 8)]


class TrackTable(C.Structure):
    _fields_ = [("num_streams", C.c_int32), ("capacity", C.c_int32), ("count", C.c_void_p),
                ("next_id", C.c_void_p), ("track_id", C.c_void_p), ("xyxy", C.c_void_p),
                ("confidence", C.c_void_p), ("class_id", C.c_void_p), ("age", C.c_void_p),
                ("time_since_update", C.c_void_p)]


class KalmanState(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("cov", C.c_void_p)]


ASSIGN_GREEDY, ASSIGN_OPTIMAL = 0, 1


class TrackOptions(C.Structure):
    _fields_ = [("track_thresh", C.c_float), ("match_thresh", C.c_float), ("track_buffer", C.c_int32),
                ("assignment", C.c_int32), ("kalman_in", C.POINTER(KalmanState)),
                ("kalman_out", C.POINTER(KalmanState)), ("cost_limit", C.c_double),
                ("assign_scratch", C.c_void_p), ("assign_scratch_bytes", C.c_size_t)]


def track_options(track_thresh, match_thresh, track_buffer, assignment="greedy", kalman_in=None, kalman_out=None,
                  assign_scratch=None):
    """rtm_track_options; ``assignment`` = "greedy" (what the reference runs without lap, tracker.py:182-194)
    or "lapjv" (its lap.lapjv branch, tracker.py:168-181: cost_limit = 1 - match_thresh in double)."""
    if assignment not in ("greedy", "lapjv"):
        raise ValueError(f"assignment must be 'greedy' or 'lapjv', not {assignment!r}")
    return TrackOptions(float(track_thresh), float(match_thresh), int(track_buffer),
                        ASSIGN_OPTIMAL if assignment == "lapjv" else ASSIGN_GREEDY,
                        C.pointer(kalman_in) if kalman_in is not None else None,
                        C.pointer(kalman_out) if kalman_out is not None else None,
                        1 - float(match_thresh),
                        assign_scratch.data_ptr() if assign_scratch is not None else None,
                        assign_scratch.numel() if assign_scratch is not None else 0)


class ZoneSet(C.Structure):
    _fields_ = [("num_streams", C.c_int32), ("num_columns", C.c_int32), ("zone_offsets", C.c_void_p),
                ("poly_offsets", C.c_void_p), ("poly_xy", C.c_void_p), ("dwell_sec", C.c_void_p),
                ("cooldown_sec", C.c_void_p), ("column", C.c_void_p)]


class ZoneState(C.Structure):
    _fields_ = [("first_seen", C.c_void_p), ("last_alert", C.c_void_p)]


class ZoneEventRec(C.Structure):
    _fields_ = [("stream", C.c_int32), ("frame_id", C.c_int32), ("track_id", C.c_int32),
                ("zone", C.c_int32), ("class_id", C.c_int32), ("cx", C.c_int32), ("cy", C.c_int32),
                ("row", C.c_int32), ("dwell", C.c_double), ("now", C.c_double), ("xyxy", C.c_float * 4)]


#: numpy view of rtm_zone_event (64 bytes)
EVENT_DTYPE = [("stream", "<i4"), ("frame_id", "<i4"), ("track_id", "<i4"), ("zone", "<i4"),
               ("class_id", "<i4"), ("cx", "<i4"), ("cy", "<i4"), ("row", "<i4"), ("dwell", "<f8"),
               ("now", "<f8"), ("xyxy", "<f4", (4,))]


class StepIO(C.Structure):
    _fields_ = [
        ("head_p3", C.c_void_p), ("head_p4", C.c_void_p), ("head_p5", C.c_void_p),
        ("head_dtype", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("scale", C.c_void_p), ("det_xyxy", C.c_void_p), ("det_conf", C.c_void_p),
        ("det_cls", C.c_void_p), ("det_anchor", C.c_void_p), ("det_keep", C.c_void_p),
        ("det_count", C.c_void_p), ("det_stride", C.c_int32), ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("table_in", C.POINTER(TrackTable)), ("table_out", C.POINTER(TrackTable)),
        ("track_thresh", C.c_float), ("match_thresh", C.c_float), ("track_buffer", C.c_int32),
        ("det_track_id", C.c_void_p), ("det_kind", C.c_void_p), ("src_row", C.c_void_p),
        ("zones", C.POINTER(ZoneSet)), ("state_in", C.POINTER(ZoneState)),
        ("state_out", C.POINTER(ZoneState)), ("now", C.c_double), ("now_per_stream", C.c_void_p),
        ("frame_id", C.c_int32), ("events", C.c_void_p), ("event_stride", C.c_int32),
        ("event_count", C.c_void_p), ("status", C.c_void_p),
        ("kalman_in", C.POINTER(KalmanState)), ("kalman_out", C.POINTER(KalmanState)),
        ("assignment", C.c_int32), ("cost_limit", C.c_double),
        ("assign_scratch", C.c_void_p), ("assign_scratch_bytes", C.c_size_t),
        ("scan_async", C.c_int32), ("heads_ready_event", C.c_void_p), ("results_alternate", C.c_int32),
    ]


class StepHostIO(C.Structure):
    _fields_ = [("host_head_p3", C.c_void_p), ("host_head_p4", C.c_void_p), ("host_head_p5", C.c_void_p),
                ("host_events", C.c_void_p), ("host_event_count", C.c_void_p),
                ("host_det_xyxy", C.c_void_p), ("host_det_conf", C.c_void_p),
                ("host_det_cls", C.c_void_p), ("host_det_track_id", C.c_void_p),
                ("host_det_count", C.c_void_p), ("host_status", C.c_void_p),
                ("wait_event", C.c_void_p), ("done_event", C.c_void_p), ("host_event_stride", C.c_int32),
                ("copy_wait_event", C.c_void_p), ("copy_done_event", C.c_void_p)]


#: every symbol include/rtmodt_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rtm_version": (C.c_int, []),
    "rtm_last_error": (C.c_char_p, []),
    "rtm_device_info": (C.c_int, [i32p, i32p, i32p]),
    "rtm_letterbox": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "rtm_letterbox_ex": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]),
    "rtm_nms_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "rtm_workspace_release": (C.c_int, [C.c_void_p]),
    "rtm_decode_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "rtm_nms_pred": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(NmsParams), C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rtm_decode_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "rtm_track_step": (C.c_int, [C.POINTER(TrackTable), C.POINTER(TrackTable), C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtm_track_step_ex": (C.c_int, [C.POINTER(TrackTable), C.POINTER(TrackTable), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(TrackOptions),
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtm_assign_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "rtm_zone_step": (C.c_int, [C.POINTER(ZoneSet), C.POINTER(TrackTable), C.c_void_p,
                                C.POINTER(ZoneState), C.POINTER(ZoneState), C.c_double, C.c_void_p,
                                C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtm_state_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "rtm_state_export": (C.c_int, [C.POINTER(TrackTable), C.POINTER(ZoneState), C.c_int32, C.POINTER(KalmanState),
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "rtm_state_import": (C.c_int, [C.POINTER(TrackTable), C.POINTER(ZoneState), C.c_int32, C.POINTER(KalmanState),
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "rtm_profile_enable": (C.c_int, [C.c_int32]),
    "rtm_profile_read": (C.c_int, [f64p, i32p]),
    "rtm_post_backbone_step": (C.c_int, [C.POINTER(StepIO), C.POINTER(NmsParams), C.c_void_p]),
    "rtm_post_backbone_step_host": (C.c_int, [C.POINTER(StepIO), C.POINTER(StepHostIO),
                                              C.POINTER(NmsParams), C.c_void_p]),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the library and type every declared symbol (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RtmError(f"{path} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                       "There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def lib():
    """The loaded library, after checking that a CUDA device is usable."""
    import torch
    if not torch.cuda.is_available():
        raise RtmError("rtmodt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return load_library()


def check(rc: int) -> None:
    if rc != 0:
        raise RtmError(f"rtm call failed ({rc}): {load_library().rtm_last_error().decode()}")


def raise_on_status(status_host, what: str = "") -> None:
    """Turn non-zero per-stream status words (host numpy / list) into an exception."""
    import numpy as np
    st = np.asarray(status_host)
    if not st.any():
        return
    bits = int(np.bitwise_or.reduce(st.reshape(-1)))
    msgs = [t for b, t in _STATUS_TEXT.items() if bits & b]
    streams = np.flatnonzero(st.reshape(-1))[:8].tolist()
    raise RtmError(f"{what}: {'; '.join(msgs)} (streams {streams})")


K_LETTERBOX, K_DECODE, K_NMS, K_TRACK, K_ZONE, K_PRED, K_POST, K_STEP, K_COUNT = 0, 1, 2, 3, 4, 5, 6, 7, 8
KERNEL_NAMES = {K_LETTERBOX: "letterbox", K_DECODE: "decode", K_NMS: "nms", K_TRACK: "track", K_ZONE: "zone", K_PRED: "pred_filter", K_POST: "post", K_STEP: "step"}


def profile_read():
    """{kernel name: (total ms, launches)} since the last read; needs rtm_profile_enable(1)."""
    ms = (C.c_double * K_COUNT)()
    n = (C.c_int32 * K_COUNT)()
    check(load_library().rtm_profile_read(ms, n))
    return {KERNEL_NAMES[k]: (ms[k], n[k]) for k in KERNEL_NAMES if n[k]}


def ptr(t) -> int:
    """Device (or pinned host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


_raw_stream = None


def cuda_stream(device_index=None) -> int:
    """Handle of torch's current CUDA stream on ``device_index`` (default: the current device).  The raw getter is a
    tenth of the cost of building a ``torch.cuda.Stream`` object (0.3 vs 3.3 us - a quarter of a step's host side)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device() if device_index is None else device_index)
    if device_index is None:
        return torch.cuda.current_stream().cuda_stream
    return torch.cuda.current_stream(device_index).cuda_stream


def dtype_code(torch_dtype) -> int:
    import torch
    try:
        return {torch.float32: RTM_F32, torch.float16: RTM_F16, torch.bfloat16: RTM_BF16}[torch_dtype]
    except KeyError:
        raise TypeError(f"unsupported tensor dtype {torch_dtype}") from None


def make_nms_params(conf=0.35, iou=0.45, max_det=100, agnostic=False, classes=None, num_classes=80):
    p = NmsParams()
    p.iou_thres, p.conf_thres, p.max_det = float(iou), float(conf), int(max_det)
    p.agnostic, p.num_classes = int(bool(agnostic)), int(num_classes)
    mask = [0] * 8
    wanted = range(num_classes) if classes is None else classes
    for c in wanted:
        c = int(c)
        if 0 <= c < 256:
            mask[c >> 5] |= 1 << (c & 31)
    for k in range(8):
        p.class_mask[k] = mask[k]
    return p
