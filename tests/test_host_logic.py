"""Host-side logic of the product (no GPU): geometry mirrors, masks, sharding, record schema."""
import os
import numpy as np
import pytest
import torch

from manual_yolo_b200 import geometry, pipeline, synth
from oracle import boxes as oboxes
from oracle import letterbox as olb


@pytest.mark.parametrize("hw", [(1200, 1920), (900, 1600), (1130, 930), (543, 770), (1194, 1919), (1034, 1700),
                                (64, 64), (100, 37), (2160, 3840)])
@pytest.mark.parametrize("new", [640, 1280, (384, 640)])
@pytest.mark.parametrize("auto", [False, True])
def test_letterbox_geometry_matches_oracle(hw, new, auto):
    a = geometry.letterbox_geometry(hw, new, auto=auto)
    b = olb.letterbox_geometry(hw, new, auto=auto)
    assert a == b


def test_scale_params_and_levels():
    gain, pad = geometry.scale_boxes_params((640, 640), (1200, 1920))
    assert gain == 640 / 1920 and pad == (0, 120)
    gain, pad = geometry.scale_boxes_params((384, 640), (900, 1600))
    assert pad == (0, 12)
    assert geometry.level_shapes(640, 640) == [(80, 80), (40, 40), (20, 20)]
    assert geometry.num_anchors(640, 640) == 8400 and geometry.num_anchors(384, 640) == 5040
    assert geometry.num_anchors(640, 544) == 7140 and geometry.num_anchors(928, 1280) == 24360


def test_class_mask_and_shards():
    assert geometry.class_mask_words([0, 31, 32, 63], 64) == [0x80000001, 0x80000001]
    assert geometry.class_mask_words(pipeline.RANK_CLASS_IDS, 64) == [
        sum(1 << c for c in pipeline.RANK_CLASS_IDS if c < 32), sum(1 << (c - 32) for c in pipeline.RANK_CLASS_IDS if c >= 32)]
    spans = [geometry.shard_range(1000, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 1000
    assert all(spans[i][1] == spans[i + 1][0] for i in range(7))
    assert [b - a for a, b in spans] == [125] * 8
    assert [geometry.shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_detections_to_records_schema():
    rows = torch.zeros((2, 3, 6))
    rows[0, 0] = torch.tensor([10.9, 20.2, 30.7, 40.1, 0.87654, 6.0])
    rows[1, 0] = torch.tensor([1.0, 2.0, 3.0, 4.0, 0.5, 63.0])
    recs = pipeline.detections_to_records(rows, torch.tensor([1, 1]), names={6: "card1_rank"}, frame_offset=7)
    assert recs[0] == {"frame": 7, "tracker_id": -1, "class_id": 6, "class_name": "card1_rank",
                       "bbox": [10, 20, 30, 40], "conf": 0.877}
    assert recs[1]["class_name"] == "class63" and recs[1]["frame"] == 8


def test_synth_generators_are_seeded_and_guard_banded():
    labels = synth.load_labels()
    assert labels["boxes"].shape == (4259, 5) and labels["img_hw"].shape == (200, 2)
    h1, gts = synth.synth_head_from_labels(2, 64, seed=3, labels=labels)
    h2, _ = synth.synth_head_from_labels(2, 64, seed=3, labels=labels)
    assert torch.equal(h1, h2) and h1.shape == (2, 128, 8400)
    score = h1[:, 64:].max(1).values.sigmoid()
    bits = score.view(torch.int32).long()
    thr = int(torch.tensor([0.25]).view(torch.int32))
    assert ((bits - thr).abs() > 16).all()
    n = (score > 0.25).sum(1)
    assert (n > 5).all() and (n < 400).all()
    d = synth.synth_head_dense(1, 80, seed=0)
    assert d.shape == (1, 144, 8400)
    assert (d[:, 64:].max(1).values.sigmoid() > 0.001).float().mean() > 0.99
    boxes, bidx = synth.synth_rois(256, 8, seed=0)
    assert boxes.shape == (256, 4) and bidx.max() < 8
    crops = [oboxes.safe_crop_box_ref((1200, 1920), *[int(v) for v in b]) for b in boxes]
    assert all(c is not None for c in crops)


def test_decoded_synthetic_head_round_trips_to_labels():
    """Label-derived heads decode (through the oracle) back to the ground-truth boxes within ~1.5 px."""
    from oracle import head as ohead
    from oracle import nms as onms
    labels = synth.load_labels()
    head, gts = synth.synth_head_from_labels(2, 64, seed=5, labels=labels)
    pred = ohead.detect_inference_ref(head, geometry.level_shapes(640, 640))
    out = onms.non_max_suppression_ref(pred, 0.25, 0.7)
    for o, gt in zip(out, gts):
        assert 0.6 * len(gt) <= o.shape[0] <= 1.5 * len(gt) + 5
        # every detection sits on a ground-truth box of its class
        for row in o[:10]:
            same = gt[gt[:, 4] == float(row[5])]
            assert same.shape[0] > 0
            assert np.abs(same[:, :4] - row[:4].numpy()).max(1).min() < 3.0


def test_ultralytics_shim_is_import_guarded():
    """Ultralytics is not installed here: the shim must say so loudly instead of half-installing."""
    import importlib.util
    from manual_yolo_b200 import ultralytics_shim
    if importlib.util.find_spec("ultralytics") is None:
        with pytest.raises(ImportError):
            ultralytics_shim.install()


def test_referenced_rows_against_cv2():
    """geometry.referenced_rows: for odd integer vertical scales cv2.resize(INTER_LINEAR) must depend on the
    named rows only (poisoning every other row leaves the output unchanged), and the staged image with a
    vertical scale of 1 must give the same bytes.  Other scales report the full range."""
    import cv2
    import numpy as np
    from manual_yolo_b200 import geometry
    rng = np.random.default_rng(3)
    for (H, W), (nh, nw) in [((1200, 1920), (400, 640)), ((300, 480), (100, 160)), ((500, 800), (100, 160))]:
        row0, step, n = geometry.referenced_rows(H, nh)
        k = H // nh
        assert (row0, step, n) == ((k - 1) // 2, k, nh)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
        poisoned = img.copy()
        keep = np.zeros(H, bool)
        keep[row0::step] = True
        poisoned[~keep] = rng.integers(0, 256, ((~keep).sum(), W, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(poisoned, (nw, nh), interpolation=cv2.INTER_LINEAR), ref)
        staged = np.ascontiguousarray(img[row0::step][:n])
        assert staged.shape[0] == nh
        assert np.array_equal(cv2.resize(staged, (nw, nh), interpolation=cv2.INTER_LINEAR), ref)
    for H, nh in [(900, 360), (1200, 600), (543, 903), (640, 640), (1130, 640)]:     # 2.5, even, up-scale, identity
        assert geometry.referenced_rows(H, nh) == (0, 1, H)


def test_slice_boxes_layout():
    """geometry.slice_boxes (SAHI get_slice_bboxes layout, pipe.py:183-194 parameters): equal-size windows inside
    the image, full coverage, the documented counts, and agreement with the oracle's independent restatement."""
    import numpy as np
    from manual_yolo_b200 import geometry
    from oracle import slicing
    for (H, W), n in [((1200, 1920), 12), ((543, 770), 2), ((640, 640), 1), ((900, 1600), 6), ((100, 100), 1)]:
        sl = geometry.slice_boxes(H, W, 640, 640, 0.2, 0.2)
        assert sl == slicing.get_slice_bboxes_ref(H, W, 640, 640, 0.2, 0.2)
        assert len(sl) == n
        sizes = {(x1 - x0, y1 - y0) for x0, y0, x1, y1 in sl}
        assert sizes == {(min(640, W), min(640, H))}
        cover = np.zeros((H, W), bool)
        for x0, y0, x1, y1 in sl:
            assert 0 <= x0 < x1 <= W and 0 <= y0 < y1 <= H
            cover[y0:y1, x0:x1] = True
        assert cover.all()
    assert geometry.slice_boxes(1200, 1920)[:3] == [[0, 0, 640, 640], [512, 0, 1152, 640], [1024, 0, 1664, 640]]
    assert geometry.slice_boxes(1200, 1920)[-1] == [1280, 560, 1920, 1200]          # shifted back inside
    import pytest
    with pytest.raises(ValueError):
        geometry.slice_boxes(100, 100, 10, 10, 1.0, 1.0)


def test_tracker_handoff_and_jsonl_emission(tmp_path):
    """N2/N4 host formats: sv.Detections-style arrays with create_clean_detections' cleaning rules
    (detect.py:253-310) and append-only JSONL with the reference's per-frame object (detect.py:590-598, 679-683)."""
    import json
    import numpy as np
    import torch
    from manual_yolo_b200 import handoff
    rows = torch.zeros((3, 4, 6))
    rows[0, 0] = torch.tensor([10.9, 20.2, 30.7, 40.99, 0.87654, 6.0])
    rows[0, 1] = torch.tensor([1.0, 2.0, 3.0, 4.0, float("nan"), float("nan")])
    rows[2, 0] = torch.tensor([5.5, 6.5, 7.5, 8.5, 0.5, 11.0])
    count = torch.tensor([2, 0, 1])
    arr = handoff.to_tracker_arrays(rows, count, tracker_ids=[[3, None], [], [float("nan")]])
    assert [a["xyxy"].shape for a in arr] == [(2, 4), (0, 4), (1, 4)]
    assert arr[0]["xyxy"].dtype == np.float32 and arr[0]["class_id"].dtype == np.int32
    assert arr[0]["class_id"].tolist() == [6, 0] and arr[0]["confidence"].tolist() == [np.float32(0.87654), 0.0]
    assert arr[0]["tracker_id"].tolist() == [3, -1] and arr[2]["tracker_id"].tolist() == [-1]
    assert handoff.to_tracker_arrays(rows, count)[0]["tracker_id"] is None
    rows[0, 1, 4:] = torch.tensor([0.25, 3.0])
    objs = handoff.frame_objects(rows, count, names={6: "card1_rank"}, frame_offset=100, timestamp=1.5)
    assert objs[0]["frame"] == 100 and objs[1]["detections"] == [] and objs[2]["frame"] == 102
    d0 = objs[0]["detections"][0]
    assert d0 == {"frame": 100, "tracker_id": -1, "class_id": 6, "class_name": "card1_rank", "bbox": [10, 20, 30, 40],
                  "conf": 0.877, "ocr_text": ""}
    assert objs[0]["detections"][1]["class_name"] == "class3"
    p = tmp_path / "det.jsonl"
    with handoff.JsonlWriter(str(p)) as w:
        w.write_batch(objs)
        first = w.bytes
        w.write_batch(handoff.frame_objects(rows, count, frame_offset=103, timestamp=2.5))
        assert w.frames == 6 and w.bytes < 2.2 * first           # appended, not re-dumped
    back = handoff.load_jsonl_as_reference_list(str(p))
    assert [o["frame"] for o in back] == [100, 101, 102, 103, 104, 105] and back[0] == json.loads(json.dumps(objs[0]))


def test_reference_helpers_golden(golden_dir):
    """Golden vectors produced by EXECUTING the reference's own helpers (tests/golden/make_rank_text_golden.py cuts
    normalize_rank_text and safe_crop out of /root/reference/detect.py and runs them unmodified): they pin the
    classifier hand-off text rules and the crop geometry the oracle (and through it K5) follows."""
    import json
    import os
    import numpy as np
    from manual_yolo_b200 import handoff
    from oracle import boxes as oboxes
    g = json.load(open(os.path.join(golden_dir, "rank_text_golden.json")))
    assert sorted(handoff.VALID_CARD_RANKS) == g["valid_card_ranks"]
    for text, want in g["normalize_rank_text"]:
        assert handoff.normalize_rank_text(text) == want, (text, want)
    H, W = g["frame_hw"]
    frame = np.zeros((H, W, 3), np.uint8)
    for x1, y1, x2, y2, pad, shape in g["safe_crop_shapes"]:
        crop = oboxes.safe_crop_ref(frame, x1, y1, x2, y2, pad=pad)
        assert (None if crop is None else list(crop.shape[:2])) == shape, (x1, y1, x2, y2, pad)
    # classify_card_rank (detect.py:115-139), executed unmodified against a stand-in rank_model: 1 309 (top-1 name,
    # confidence, detection class name) -> text decisions, thresholds probed on both sides in float32
    assert len(g["classify_card_rank"]) > 1000 and g["classify_card_rank_empty"] == ["", ""]
    for pred, conf, cname, want in g["classify_card_rank"]:
        assert handoff.rank_text_from_top1(pred, conf, cname) == want, (pred, conf, cname, want)
    assert handoff.rank_text_from_top1("k", 0.41, "card1_rank") == "K"
    assert handoff.rank_text_from_top1("k", 0.39, "card1_rank") == ""
    assert handoff.rank_text_from_top1("10", 0.21, "turn_rank") == "10" and handoff.rank_text_from_top1("10", 0.19, "river_rank") == ""
    assert handoff.rank_text_from_top1("joker", 0.9, "flop1_rank") == "JOKER"


def test_pipe_records_golden(golden_dir):
    """Row a11 consumer: handoff.to_pipe_records == the reference's own parse_ultralytics_results (pipe.py:100-134),
    whose outputs on seeded boxes were recorded by executing it (tests/golden/make_pipe_records_golden.py)."""
    import json
    import os
    import torch
    from manual_yolo_b200 import handoff
    g = json.load(open(os.path.join(golden_dir, "pipe_records_golden.json")))
    names = {int(k): v for k, v in g["names"].items()}
    for case in g["cases"]:
        n = len(case["rows"])
        rows = torch.zeros((1, max(n, 1), 6))
        if n:
            rows[0, :n] = torch.tensor(case["rows"], dtype=torch.float32)
        got = handoff.to_pipe_records(rows, torch.tensor([n]), names, case["image_shape"])[0]
        assert got == case["records"]


def test_clean_detections_golden(golden_dir):
    """Tracker hand-off cleaning rules == the reference's own create_clean_detections (detect.py:253-310), whose
    outputs were recorded by executing it with a stand-in for sv.Detections (make_clean_detections_golden.py).
    Cases expressible in the padded device layout (NaN class / confidence, None / NaN tracker ids, empty frame)."""
    import json
    import os
    import numpy as np
    import torch
    from manual_yolo_b200 import handoff
    g = json.load(open(os.path.join(golden_dir, "clean_detections_golden.json")))

    def f(v):
        return float("nan") if v in ("nan", None) else float(v)
    for case in g["cases"]:
        cin = case["in"]
        n = len(cin["xyxy"])
        if cin["class_id"] is None or cin["confidence"] is None:
            continue                                   # "argument omitted" has no counterpart: the device rows always carry both
        rows = torch.zeros((1, max(n, 1), 6))
        for i in range(n):
            rows[0, i] = torch.tensor(cin["xyxy"][i] + [f(cin["confidence"][i]), f(cin["class_id"][i])])
        tid = None if cin["tracker_id"] is None else [[None if v is None else (float("nan") if v == "nan" else v) for v in cin["tracker_id"]]]
        got = handoff.to_tracker_arrays(rows, torch.tensor([n]), tracker_ids=tid)[0]
        if case["empty"]:
            assert got["xyxy"].shape == (0, 4)
            continue
        assert got["xyxy"].tolist() == case["xyxy"] and str(got["xyxy"].dtype) == case["xyxy_dtype"]
        assert got["class_id"].tolist() == case["class_id"] and str(got["class_id"].dtype) == case["class_id_dtype"]
        assert got["confidence"].tolist() == case["confidence"] and str(got["confidence"].dtype) == case["confidence_dtype"]
        assert (None if got["tracker_id"] is None else got["tracker_id"].tolist()) == case["tracker_id"]


# ---- product-side rank classifier loader (detect.py:21 without Ultralytics) -------------------------------------------
def _kat_npz(golden_dir):
    import os
    return np.load(os.path.join(golden_dir, "rank_classifier_kat.npz"))


def test_product_classifier_from_arrays_matches_oracle(golden_dir):
    from manual_yolo_b200 import classifier as pc
    from oracle import classifier as oc
    z = _kat_npz(golden_dir)
    clf = pc.rank_classifier_from_arrays({k[2:]: z[k] for k in z.files if k.startswith("w:")}, oc.NAMES)
    x = torch.from_numpy(z["roi_u8"]).float().div(255)
    logits = clf.forward_logits(x)
    assert torch.equal(logits, oc.forward_logits(oc.state_dict_from_npz(z), x))      # independent code, same torch ops
    top1, conf = clf.predict(x)
    assert int((top1.numpy() == z["labels"]).sum()) == 63                            # runs/rank_classifier/results.csv:21
    assert float(conf.min()) > 0.0 and float(conf.max()) <= 1.0


@pytest.mark.reference
@pytest.mark.skipif(not os.path.exists("/root/reference/rank_classifier.pt"), reason="reference checkpoint not on this box")
def test_product_classifier_loads_reference_checkpoint(golden_dir):
    """rank_classifier.pt (detect.py:21) through the restricted unpickler: no Ultralytics, no oracle."""
    import pickle
    from manual_yolo_b200 import classifier as pc
    clf = pc.load_rank_classifier("/root/reference/rank_classifier.pt")
    assert clf.names == {0: "10", 1: "2", 2: "3", 3: "4", 4: "5", 5: "6", 6: "7", 7: "8", 8: "9", 9: "A", 10: "J", 11: "K", 12: "Q"}
    assert clf.imgsz == 64 and clf.bn_eps == 1e-5
    z = _kat_npz(golden_dir)
    x = torch.from_numpy(z["roi_u8"]).float().div(255)
    assert int((clf.predict(x)[0].numpy() == z["labels"]).sum()) == 63
    # a pickle that references anything outside torch / containers / ultralytics shells is refused
    import io

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("true",))
    with pytest.raises(pickle.UnpicklingError):
        pc._RestrictedUnpickler(io.BytesIO(pickle.dumps(Evil()))).load()
    # ... including functions that live in allowed packages (only layer / transform classes and rebuild helpers resolve)
    for mod, name in (("torch", "load"), ("torch.hub", "load"), ("builtins", "eval"), ("builtins", "getattr"),
                      ("torch.nn.modules.module", "_addindent"), ("torchvision.transforms.functional", "to_pil_image"),
                      ("torch.storage", "_load_from_bytes"), ("os", "system")):
        with pytest.raises(pickle.UnpicklingError):
            pc._RestrictedUnpickler(io.BytesIO(b"")).find_class(mod, name)
    assert pc._RestrictedUnpickler(io.BytesIO(b"")).find_class("torch.nn.modules.conv", "Conv2d") is torch.nn.Conv2d


def test_div255_two_constant_form_is_exact():
    """K5's table-free v/255 (csrc/roi.cu top_byte_div255): RN(v*c_hi + RN(v*c_lo)) == RN(v/255) for every v in 0..255,
    evaluated with exact rational arithmetic (one rounding per fp32 op, as FMUL / FFMA round)."""
    from fractions import Fraction as Fr

    def rn32(fr):
        c = np.float32(float(fr))
        best = None
        for cand in (np.nextafter(c, np.float32(-np.inf)), c, np.nextafter(c, np.float32(np.inf))):
            d = abs(Fr(float(cand)) - fr)
            even = (int(np.float32(cand).view(np.uint32)) & 1) == 0
            if best is None or d < best[0] or (d == best[0] and even):
                best = (d, np.float32(cand))
        return best[1]
    c_hi = np.uint32(0x3B808081).view(np.float32)
    c_lo = np.uint32(0xAF7F00BF).view(np.float32)
    for v in range(256):
        t = rn32(Fr(v) * Fr(float(c_lo)))
        r = rn32(Fr(v) * Fr(float(c_hi)) + Fr(float(t)))
        assert r == np.float32(v) / np.float32(255), v


def test_round2_golden_fixtures_are_consistent(golden_dir):
    """The committed round-2 fixtures: test2.png is the 1600x900 BGRA capture; 22 frames x 2 letterbox modes; 579 crops,
    67 of them the validation split with 63 oracle hits (results.csv:21)."""
    import json
    import os
    import cv2
    gold = json.load(open(os.path.join(golden_dir, "letterbox_golden_r2.json")))
    assert len(gold) == 44
    assert cv2.imread(os.path.join(golden_dir, "frames", "test2.png"), cv2.IMREAD_UNCHANGED).shape == (900, 1600, 4)
    assert gold["frames/test2.png|auto=1"]["shape"] == [384, 640, 3] and gold["frames/test2.png|auto=0"]["shape"] == [640, 640, 3]
    sizes = sorted({tuple(v["src_hw"]) for v in gold.values()})
    assert sizes == [(900, 1600), (1034, 1700), (1200, 1920)]
    # the recorded hashes are what the real cv2 leaves give on this box too
    import hashlib
    from oracle import letterbox as olb2
    for key in ("frames/test2.png|auto=0", "frames/test2.png|auto=1", "frames_r2/f1700x1034_00.jpg|auto=1"):
        name, auto = key.split("|auto=")
        im = cv2.imread(os.path.join(golden_dir, name))
        lb = olb2.letterbox_ref(im, (640, 640), auto=bool(int(auto)))
        assert hashlib.sha256(lb.tobytes()).hexdigest() == gold[key]["sha256"]
    z = np.load(os.path.join(golden_dir, "rank_crops_all.npz"))
    v = z["split"] == 1
    assert len(z["labels"]) == 579 and int(v.sum()) == 67
    assert int((z["oracle_top1"][v] == z["labels"][v]).sum()) == 63


def test_oracle_greedy_nmm_semantics():
    """oracle/slicing.py restatement of SAHI's GREEDYNMM on hand-checkable cases: IOS merges a small box contained in a
    large one (IoU would not), merged box = union, score = max, other classes untouched, >= in the greedy match but a
    strict > in has_match."""
    from oracle import slicing
    p = torch.tensor([[0., 0., 100., 100., 0.9, 1.], [10., 10., 40., 40., 0.8, 1.],      # contained: IOS = 1, IoU = 0.09
                      [90., 90., 130., 130., 0.7, 1.],                                   # IOS = 100/1600 < 0.5: stays
                      [0., 0., 100., 100., 0.6, 2.],                                     # same box, other class: stays
                      [200., 200., 240., 240., 0.5, 1.], [220., 200., 260., 240., 0.4, 1.]])  # IOS = 0.5 exactly
    out, kept = slicing.greedy_nmm_postprocess_ref(p, "IOS", 0.5)
    got = {tuple(r) for r in out.tolist()}
    exp = {(0., 0., 100., 100., 0.9, 1.), (90., 90., 130., 130., 0.7, 1.), (0., 0., 100., 100., 0.6, 2.),
           # IOS == threshold: matched by greedy_nmm (>=, consumed) but NOT merged by has_match (>): the box disappears
           (200., 200., 240., 240., 0.5, 1.)}
    assert {tuple(round(v, 4) for v in r) for r in got} == {tuple(round(v, 4) for v in r) for r in exp}
    out_iou, _ = slicing.greedy_nmm_postprocess_ref(p, "IOU", 0.5)
    assert out_iou.shape[0] == 6                                                         # nothing reaches IoU 0.5
    # a chain: B overlaps A, C overlaps the union of A and B more than A alone -> merged box grows as it merges
    q = torch.tensor([[0., 0., 50., 50., 0.9, 0.], [20., 0., 80., 50., 0.8, 0.], [45., 0., 70., 50., 0.7, 0.]])
    out, _ = slicing.greedy_nmm_postprocess_ref(q, "IOS", 0.5, class_agnostic=True)
    exp2 = torch.tensor([[0., 0., 80., 50., 0.9, 0.], [45., 0., 70., 50., 0.7, 0.]])            # C only overlaps the grown box
    assert torch.equal(torch.tensor(sorted(out.tolist())), exp2)


def test_oracle_bytetrack_restatement_tracks_a_simple_scene():
    """oracle/bytetrack.py sanity (no GPU): two objects crossing the frame keep their ids, a 10-frame occlusion is
    bridged (lost -> re-activated, same id), a one-frame false positive never gets a confirmed id twice."""
    from oracle import bytetrack as obt
    trk = obt.ByteTrackRef()
    ids = []
    for f in range(40):
        boxes = [[100 + 5 * f, 100, 160 + 5 * f, 180, 0.9]]
        if not (15 <= f < 25):
            boxes.append([800 - 4 * f, 300, 880 - 4 * f, 420, 0.8])
        if f == 30:
            boxes.append([1200, 600, 1250, 660, 0.7])
        b = np.asarray(boxes, np.float32)
        ids.append(trk.update_with_detections(b[:, :4], b[:, 4]).tolist())
    assert all(r[0] == ids[0][0] and r[0] > 0 for r in ids)               # object A: one id throughout
    b_ids = {r[1] for f, r in enumerate(ids) if not (15 <= f < 25) and len(r) > 1 and f != 30}
    assert len(b_ids) == 1 and b_ids != {ids[0][0]}                       # object B: same id before and after the gap
    assert trk.max_time_lost == 30 and trk.det_thresh == 0.35
