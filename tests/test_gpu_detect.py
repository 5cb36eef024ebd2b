"""Letterbox, head decode and NMS kernels against the oracle (cv2 / torch CPU / torchvision)."""

import ctypes as C

import numpy as np
import pytest

from oracle import detect_ref

pytestmark = pytest.mark.gpu

WANTED = [0, 1, 2, 3, 5, 7]


def run_nms_pred(pkg, pred, conf=0.35, iou=0.45, classes=None, agnostic=False, max_det=100, scale=None):
    import torch
    lib = pkg._lib.lib()
    dev = torch.device("cuda:0")
    B, ch, A = pred.shape
    p = pkg._lib.make_nms_params(conf, iou, max_det, agnostic, classes, ch - 4)
    d_pred = torch.from_numpy(np.ascontiguousarray(pred, np.float32)).to(dev)
    out = dict(xyxy=torch.zeros(B, max_det, 4, device=dev), conf=torch.zeros(B, max_det, device=dev),
               cls=torch.zeros(B, max_det, dtype=torch.int32, device=dev), anchor=torch.zeros(B, max_det, dtype=torch.int32, device=dev),
               keep=torch.zeros(B, max_det, dtype=torch.int32, device=dev), count=torch.zeros(B, dtype=torch.int32, device=dev))
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    ws = torch.zeros(lib.rtm_nms_workspace_bytes(B, A), dtype=torch.uint8, device=dev)
    d_scale = None if scale is None else torch.tensor(scale, dtype=torch.float32, device=dev)
    pkg._lib.check(lib.rtm_nms_pred(d_pred.data_ptr(), B, A, C.byref(p), pkg._lib.ptr(d_scale), out["xyxy"].data_ptr(),
                                    out["conf"].data_ptr(), out["cls"].data_ptr(), out["anchor"].data_ptr(), out["keep"].data_ptr(),
                                    out["count"].data_ptr(), max_det, status.data_ptr(), ws.data_ptr(), ws.numel(),
                                    pkg._lib.cuda_stream()))
    assert not status.cpu().numpy().any()
    return {k: v.cpu().numpy() for k, v in out.items()}


def random_pred(rng, B, A, nc=80, n_obj=25, dup=6, tie=False, bg=0.01, fixed_cls=None, size=(20, 200)):
    """(B, 4+nc, A) prediction tensor: low background, clusters of near-duplicate boxes."""
    pred = np.zeros((B, 4 + nc, A), np.float32)
    pred[:, 4:] = rng.uniform(0, bg, (B, nc, A))
    pred[:, 0] = rng.uniform(0, 640, (B, A)); pred[:, 1] = rng.uniform(0, 640, (B, A))
    pred[:, 2] = rng.uniform(4, 200, (B, A)); pred[:, 3] = rng.uniform(4, 200, (B, A))
    for b in range(B):
        for _ in range(n_obj):
            c = int(rng.integers(0, nc)) if fixed_cls is None else int(rng.choice(fixed_cls))
            base = np.array([rng.uniform(50, 590), rng.uniform(50, 590), rng.uniform(*size), rng.uniform(*size)])
            for a in rng.choice(A, dup, replace=False):
                pred[b, :4, a] = base + rng.normal(0, 2.0, 4)
                s = rng.uniform(0.2, 0.95)
                if tie:
                    s = float(rng.choice([0.5, 0.75, 0.9]))
                pred[b, 4 + c, a] = s
                if rng.uniform() < 0.2:                                  # second class with the same score: first arg-max wins
                    pred[b, 4 + (c + 3) % nc, a] = s
    return pred


def check_against_oracle(pkg, pred, **kw):
    got = run_nms_pred(pkg, pred, **{k: v for k, v in kw.items() if k != "src_hw"}, scale=None)
    ref = detect_ref.non_max_suppression(pred, kw.get("conf", 0.35), kw.get("iou", 0.45), kw.get("classes"),
                                         kw.get("agnostic", False), kw.get("max_det", 100))
    for b, (dets, keep, anchor) in enumerate(ref):
        n = got["count"][b]
        assert n == len(keep), f"image {b}: {n} kept vs {len(keep)}"
        np.testing.assert_array_equal(got["keep"][b, :n], keep.numpy())          # torchvision's indices, bit-exact
        np.testing.assert_array_equal(got["anchor"][b, :n], anchor.numpy())
        np.testing.assert_array_equal(got["cls"][b, :n], dets[:, 5].numpy().astype(np.int32))
        np.testing.assert_array_equal(got["conf"][b, :n], dets[:, 4].numpy())
        np.testing.assert_array_equal(got["xyxy"][b, :n], dets[:, :4].numpy())
    return got


@pytest.mark.parametrize("case", ["default", "ties", "agnostic", "noclassfilter", "maxdet", "iou08", "many", "crowd", "crowd_maxdet",
                                  "two_big_classes", "all_classes", "huge_boxes", "huge_boxes_ties", "crowd_ties"])
def test_nms_matches_torchvision_bit_exact(pkg, case):
    rng = np.random.default_rng(hash(case) % 2**32)
    kw = dict(conf=0.35, iou=0.45, classes=WANTED, agnostic=False, max_det=100)
    pred_kw = dict(B=4, A=8400)
    if case == "ties":
        pred_kw.update(tie=True)
    if case == "agnostic":
        kw.update(agnostic=True, classes=None)
    if case == "noclassfilter":
        kw.update(classes=None)
    if case == "maxdet":
        kw.update(max_det=7, classes=None)
        pred_kw.update(n_obj=40)
    if case == "iou08":
        kw.update(iou=0.8, classes=None)
    if case == "many":                                                         # > 2048 candidates: global-memory path
        kw.update(classes=None, conf=0.05)
        pred_kw.update(B=2, n_obj=60, dup=50, bg=0.01)
    if case in ("crowd", "crowd_maxdet", "crowd_ties"):                        # one class, segment far longer than a warp handles
        kw.update(classes=None, max_det=300 if case == "crowd" else 9)
        pred_kw.update(B=3, n_obj=150, dup=10, fixed_cls=[0], size=(15, 60), tie=case == "crowd_ties")
    if case == "two_big_classes":                                              # long segments + short ones in one stream
        kw.update(classes=None)
        pred_kw.update(B=3, n_obj=120, dup=8, fixed_cls=[0] * 10 + [2] * 10 + [5, 7, 11, 40, 79], size=(15, 80))
    if case == "all_classes":
        kw.update(classes=None, max_det=100)
        pred_kw.update(B=3, n_obj=300, dup=4, size=(15, 80))
    if case in ("huge_boxes", "huge_boxes_ties"):                              # boxes wider than the class offset: classes interact
        kw.update(classes=None, iou=0.3)
        pred_kw.update(B=3, n_obj=60, dup=6, size=(3000, 12000), tie=case.endswith("ties"))
    pred = random_pred(rng, **pred_kw)
    if case == "many":
        pred[:, 4:, ::3] = np.maximum(pred[:, 4:, ::3], rng.uniform(0.0, 0.12, pred[:, 4:, ::3].shape).astype(np.float32) *
                                      (rng.uniform(size=pred[:, 4:, ::3].shape) < 0.02))
    got = check_against_oracle(pkg, pred, **kw)
    assert got["count"].sum() > 0


def test_nms_exact_threshold_and_zero_area(pkg):
    """IoU exactly 0.8 with thr 0.8: suppressed (float32 0.8 > double 0.8); zero-area boxes never are."""
    A, nc = 64, 4
    pred = np.zeros((1, 4 + nc, A), np.float32)
    pred[0, :4, 0] = [5, 5, 10, 10]; pred[0, 4, 0] = 0.9
    pred[0, :4, 1] = [5, 4, 10, 8]; pred[0, 4, 1] = 0.8           # IoU with box 0 = 0.8
    pred[0, :4, 2] = [30, 30, 0, 0]; pred[0, 5, 2] = 0.7          # zero area
    pred[0, :4, 3] = [30, 30, 0, 0]; pred[0, 5, 3] = 0.6
    pred[0, :4, 9] = [5, 5, 10, 10]; pred[0, 6, 9] = 0.9          # other class: offset keeps it
    for iou in (0.8, 0.45, float(np.float32(0.8))):
        check_against_oracle(pkg, pred, conf=0.35, iou=iou, classes=None)


def test_scale_boxes_matches_oracle(pkg):
    rng = np.random.default_rng(5)
    pred = random_pred(rng, B=2, A=8400)
    for src_hw in [(1080, 1920), (720, 1280), (480, 640), (1000, 700)]:
        gain, px, py = pkg.synth.scale_params(src_hw)
        got = run_nms_pred(pkg, pred, classes=None, scale=[[gain, px, py, src_hw[1], src_hw[0]]] * 2)
        ref = detect_ref.non_max_suppression(pred, classes=None)
        for b, (dets, keep, anchor) in enumerate(ref):
            exp = detect_ref.scale_boxes((640, 640), dets[:, :4], src_hw).numpy()
            np.testing.assert_array_equal(got["xyxy"][b, :got["count"][b]], exp)


def make_heads(pkg, B, seed, n_obj=30, dtype=np.float32):
    rng = np.random.default_rng(seed)
    levels = [[], [], []]
    for b in range(B):
        wh = np.stack([rng.uniform(30, 160, n_obj), rng.uniform(40, 220, n_obj)], 1)
        c = np.stack([rng.uniform(80, 560, n_obj), rng.uniform(80, 560, n_obj)], 1)
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1)
        cls = rng.choice(np.asarray(WANTED), n_obj)
        for l, t in enumerate(pkg.synth.plant_head(rng, boxes, cls)):
            levels[l].append(t)
    return [np.stack(l).astype(np.float32) for l in levels]


def run_decode_nms(pkg, heads_t, conf=0.35, iou=0.45, classes=WANTED, max_det=100, src_hw=(1080, 1920)):
    import torch
    lib = pkg._lib.lib()
    dev = heads_t[0].device
    B = heads_t[0].shape[0]
    p = pkg._lib.make_nms_params(conf, iou, max_det, False, classes, 80)
    gain, px, py = pkg.synth.scale_params(src_hw)
    scale = torch.tensor([[gain, px, py, src_hw[1], src_hw[0]]] * B, dtype=torch.float32, device=dev)
    out = dict(xyxy=torch.zeros(B, max_det, 4, device=dev), conf=torch.zeros(B, max_det, device=dev),
               cls=torch.zeros(B, max_det, dtype=torch.int32, device=dev), anchor=torch.zeros(B, max_det, dtype=torch.int32, device=dev),
               keep=torch.zeros(B, max_det, dtype=torch.int32, device=dev), count=torch.zeros(B, dtype=torch.int32, device=dev))
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    ws = torch.zeros(lib.rtm_nms_workspace_bytes(B, 8400), dtype=torch.uint8, device=dev)
    pkg._lib.check(lib.rtm_decode_nms(heads_t[0].data_ptr(), heads_t[1].data_ptr(), heads_t[2].data_ptr(),
                                      pkg._lib.dtype_code(heads_t[0].dtype), B, 640, 640, C.byref(p), scale.data_ptr(),
                                      out["xyxy"].data_ptr(), out["conf"].data_ptr(), out["cls"].data_ptr(), out["anchor"].data_ptr(),
                                      out["keep"].data_ptr(), out["count"].data_ptr(), max_det, status.data_ptr(), ws.data_ptr(),
                                      ws.numel(), pkg._lib.cuda_stream()))
    assert not status.cpu().numpy().any()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_decode_head_within_tolerance(pkg, dtype):
    """D1 (DFL + dist2bbox + sigmoid) vs torch CPU float32: 1e-4 relative (exp differs by ulps)."""
    import torch
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    heads = [torch.from_numpy(h).to(tdt) for h in make_heads(pkg, 2, seed=3)]
    ref = detect_ref.decode_head([h.float() for h in heads]).numpy()
    dev = torch.device("cuda:0")
    d = [h.to(dev).contiguous() for h in heads]
    pred = torch.zeros(2, 84, 8400, device=dev)
    lib = pkg._lib.lib()
    pkg._lib.check(lib.rtm_decode_head(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), pkg._lib.dtype_code(tdt), 2, 640, 640, 80,
                                       pred.data_ptr(), pkg._lib.cuda_stream()))
    got = pred.cpu().numpy()
    np.testing.assert_allclose(got[:, :4], ref[:, :4], rtol=1e-4, atol=1e-3)     # pixels
    np.testing.assert_allclose(got[:, 4:], ref[:, 4:], rtol=1e-4, atol=1e-7)     # probabilities


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_fused_decode_nms_matches_oracle_pipeline(pkg, dtype):
    """Head tensors -> detections: same anchors / classes kept as the oracle chain, boxes and
    scores within 1e-4 relative.  (Index flips could only come from ulp-level score ties.)"""
    import torch
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    heads = [torch.from_numpy(h).to(tdt) for h in make_heads(pkg, 4, seed=8)]
    ref = detect_ref.detect_post([h.float() for h in heads], (1080, 1920), classes=WANTED)
    got = run_decode_nms(pkg, [h.to("cuda:0").contiguous() for h in heads])
    flips = 0
    for b, r in enumerate(ref):
        n = got["count"][b]
        assert n == len(r["conf"]) and n > 5
        flips += int((got["anchor"][b, :n] != r["anchor"]).sum())
        if flips == 0:
            np.testing.assert_array_equal(got["keep"][b, :n], r["keep"])
            np.testing.assert_array_equal(got["cls"][b, :n], r["cls"])
            np.testing.assert_allclose(got["conf"][b, :n], r["conf"], rtol=1e-4)
            np.testing.assert_allclose(got["xyxy"][b, :n], r["xyxy"], rtol=1e-4, atol=1e-2)
    assert flips == 0, f"{flips} index flips against the oracle"


def test_fused_decode_equals_decode_then_nms_on_device(pkg):
    """Stage isolation: rtm_decode_nms == rtm_nms_pred(rtm_decode_head(x)) bit for bit."""
    import torch
    dev = torch.device("cuda:0")
    heads = [torch.from_numpy(h).to(dev) for h in make_heads(pkg, 3, seed=21)]
    fused = run_decode_nms(pkg, heads, classes=None)
    pred = torch.zeros(3, 84, 8400, device=dev)
    lib = pkg._lib.lib()
    pkg._lib.check(lib.rtm_decode_head(heads[0].data_ptr(), heads[1].data_ptr(), heads[2].data_ptr(), 0, 3, 640, 640, 80,
                                       pred.data_ptr(), pkg._lib.cuda_stream()))
    gain, px, py = pkg.synth.scale_params((1080, 1920))
    staged = run_nms_pred(pkg, pred.cpu().numpy(), classes=None, scale=[[gain, px, py, 1920, 1080]] * 3)
    np.testing.assert_array_equal(fused["count"], staged["count"])
    for b in range(3):
        n = fused["count"][b]
        np.testing.assert_array_equal(fused["anchor"][b, :n], staged["anchor"][b, :n])
        np.testing.assert_allclose(fused["conf"][b, :n], staged["conf"][b, :n], rtol=1e-6)
        np.testing.assert_allclose(fused["xyxy"][b, :n], staged["xyxy"][b, :n], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("src_hw", [(1080, 1920), (720, 1280), (480, 640), (375, 500), (64, 48), (640, 640), (1200, 800)])
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
def test_letterbox_matches_cv2_bit_exact(pkg, src_hw, dtype):
    import torch
    rng = np.random.default_rng(src_hw[0] * 7 + src_hw[1])
    B = 3
    frames = rng.integers(0, 256, (B, *src_hw, 3), dtype=np.uint8)
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    dev = torch.device("cuda:0")
    d_frames = torch.from_numpy(frames).to(dev)
    out = torch.zeros(B, 3, 640, 640, dtype=tdt, device=dev)
    lib = pkg._lib.lib()
    pkg._lib.check(lib.rtm_letterbox(d_frames.data_ptr(), B, src_hw[0], src_hw[1], src_hw[1] * 3, src_hw[0] * src_hw[1] * 3,
                                     out.data_ptr(), pkg._lib.dtype_code(tdt), 640, 640, pkg._lib.cuda_stream()))
    got = out.cpu()
    for b in range(B):
        ref = detect_ref.preprocess(detect_ref.letterbox(frames[b]), dtype)
        assert torch.equal(got[b], ref), f"frame {b}: {(got[b].float() - ref.float()).abs().max()}"


def test_detector_facade_end_to_end(pkg):
    """Detector.detect(frame): letterbox -> network -> decode -> NMS -> rescale -> Detections.
    The network is a stub returning planted head tensors, so that the expected detections do not
    depend on ulp-level score ties of a random-init conv stack."""
    import torch
    from rtmodt_b200.detection.yolov8s import YOLOv8s

    heads = [torch.from_numpy(h) for h in make_heads(pkg, 1, seed=33, n_obj=20)]

    class Planted(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.seen = None

        def forward(self, x):
            self.seen = x
            return [h.to(x.device).to(x.dtype).expand(x.shape[0], -1, -1, -1).contiguous() for h in heads]

    rng = np.random.default_rng(0)
    frame = pkg.synth.synthetic_frame(rng, 720, 1280)
    for half, tdt in ((False, torch.float32), (True, torch.bfloat16)):
        stub = Planted()
        det = pkg.Detector(None, model=stub, half=half, classes=WANTED, warmup=True)
        out = det.detect(frame)
        assert isinstance(out, pkg.Detections) and out.xyxy.dtype == np.float32 and out.class_id.dtype == np.int32
        # the network saw exactly what ultralytics' preprocess would have produced
        exp_in = detect_ref.preprocess(detect_ref.letterbox(frame), "bf16" if half else "f32")
        assert torch.equal(stub.seen[0].cpu(), exp_in)
        ref = detect_ref.detect_post([h.to(tdt).float() for h in heads], (720, 1280), classes=WANTED)[0]
        assert len(out) == len(ref["conf"]) > 5
        np.testing.assert_array_equal(out.class_id, ref["cls"])
        np.testing.assert_allclose(out.confidence, ref["conf"], rtol=1e-4)
        np.testing.assert_allclose(out.xyxy, ref["xyxy"], rtol=1e-4, atol=1e-2)
        assert out.class_names == [det.names[int(c)] for c in out.class_id]
        kept = out.filter_classes([0, 2])
        assert set(kept.class_id.tolist()) <= {0, 2} and len(kept.class_names) == len(kept)
        two = det.detect_batch(np.stack([frame, frame]))
        assert len(two) == 2 and np.array_equal(two[0].xyxy, two[1].xyxy) and np.array_equal(two[0].xyxy, out.xyxy)
    with pytest.raises(FileNotFoundError):
        pkg.Detector("weights/does_not_exist.pt", fallback_model="weights/neither.pt")


def test_detector_runs_the_yolov8s_stand_in(pkg, tmp_path):
    """The same-shape YOLOv8s module (random init, bf16) goes through the whole Detector path; with
    the stock Detect bias nothing scores above 0.35 (last hard part)."""
    import torch
    from rtmodt_b200.detection.yolov8s import YOLOv8s
    torch.manual_seed(0)
    net = YOLOv8s()
    path = tmp_path / "yolov8s_random.pt"
    torch.save(net.state_dict(), path)
    det = pkg.Detector(str(path), half=True, warmup=False)
    frame = pkg.synth.synthetic_frame(np.random.default_rng(1), 1080, 1920)
    empty = det.detect(frame)
    assert len(empty) == 0 and empty.xyxy.shape == (0, 4) and empty.class_id.dtype == np.int32
    for seq in det.model.detect.cv3:                       # raise the class bias: now it fires, capped at max_det
        seq[-1].bias.data[:] = 0.5
    busy = det.detect(frame)
    assert 0 < len(busy) <= 100 and busy.xyxy.min() >= 0 and busy.xyxy[:, [0, 2]].max() <= 1920


@pytest.mark.parametrize("src_hw", [(1080, 1920), (720, 1280), (500, 375), (333, 500)])
def test_letterbox_auto_rectangle_matches_cv2_bit_exact(pkg, src_hw):
    """LetterBox(auto=True), the variant ultralytics uses for .pt models: padding only to the next
    multiple of 32 (1080p -> 384 x 640).  rtm_letterbox_ex takes the geometry from the caller."""
    import torch
    rng = np.random.default_rng(src_hw[0] + 3 * src_hw[1])
    frames = rng.integers(0, 256, (2, *src_hw, 3), dtype=np.uint8)
    (new_w, new_h), (top, bottom, left, right) = detect_ref.letterbox_geometry(src_hw, (640, 640), auto=True)
    H, W = new_h + top + bottom, new_w + left + right
    assert H % 32 == 0 and W % 32 == 0 and (H < 640 or W < 640 or src_hw[0] == src_hw[1])
    dev = torch.device("cuda:0")
    d_frames = torch.from_numpy(frames).to(dev)
    out = torch.zeros(2, 3, H, W, dtype=torch.bfloat16, device=dev)
    lib = pkg._lib.lib()
    pkg._lib.check(lib.rtm_letterbox_ex(d_frames.data_ptr(), 2, src_hw[0], src_hw[1], src_hw[1] * 3, src_hw[0] * src_hw[1] * 3,
                                        out.data_ptr(), pkg._lib.RTM_BF16, H, W, new_h, new_w, top, left, pkg._lib.cuda_stream()))
    for b in range(2):
        ref = detect_ref.preprocess(detect_ref.letterbox(frames[b], (640, 640), auto=True), "bf16")
        assert tuple(ref.shape) == (3, H, W)
        assert torch.equal(out[b].cpu(), ref)


def test_detector_auto_rectangle_end_to_end(pkg):
    """Detector(auto=True) on a 1080p frame: 384 x 640 network input, 5040 anchors, detections
    rescaled with the rectangle's padding."""
    import torch
    rng = np.random.default_rng(2)
    n_obj = 12
    wh = np.stack([rng.uniform(40, 120, n_obj), rng.uniform(40, 120, n_obj)], 1)
    c = np.stack([rng.uniform(80, 560, n_obj), rng.uniform(70, 310, n_obj)], 1)
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1)
    cls = rng.choice(np.asarray(WANTED), n_obj)
    heads = [torch.from_numpy(t[None]) for t in pkg.synth.plant_head(rng, boxes, cls, imgsz=(384, 640), distractor_frac=0.0)]
    assert [tuple(h.shape[2:]) for h in heads] == [(48, 80), (24, 40), (12, 20)]

    class Planted(torch.nn.Module):
        seen = None

        def forward(self, x):
            Planted.seen = x
            return [h.to(x.device).to(x.dtype).contiguous() for h in heads]

    frame = pkg.synth.synthetic_frame(rng, 1080, 1920)
    det = pkg.Detector(None, model=Planted(), half=True, classes=WANTED, warmup=False, auto=True)
    out = det.detect(frame)
    assert tuple(Planted.seen.shape) == (1, 3, 384, 640)
    assert torch.equal(Planted.seen[0].cpu(), detect_ref.preprocess(detect_ref.letterbox(frame, (640, 640), auto=True), "bf16"))
    ref = detect_ref.detect_post([h.to(torch.bfloat16).float() for h in heads], (1080, 1920), imgsz=(384, 640), classes=WANTED)[0]
    assert len(out) == len(ref["conf"]) > 5
    np.testing.assert_array_equal(out.class_id, ref["cls"])
    np.testing.assert_allclose(out.xyxy, ref["xyxy"], rtol=1e-4, atol=1e-2)
