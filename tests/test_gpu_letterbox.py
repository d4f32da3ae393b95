"""K1 parity: CUDA letterbox(+normalise) vs the real-cv2 oracle, bit-exact."""
import hashlib
import json
import os

import cv2
import numpy as np
import pytest
import torch

import manual_yolo_b200 as m
from oracle import letterbox as olb

pytestmark = pytest.mark.gpu


def _frames(shape, seed, B=1):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (B, shape[0], shape[1], 3), dtype=np.uint8)


def test_div255_exhaustive(cuda_dev):
    """u8_div255 (reciprocal + Newton residual) == torch's true fp32 division for all 256 values."""
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)[None]
    out = m.preprocess(torch.from_numpy(img).to(cuda_dev), (16, 16))
    ref = torch.arange(256, dtype=torch.float32).div(255).view(16, 16)
    for c in range(3):
        assert torch.equal(out[0, c].cpu(), ref)


@pytest.mark.parametrize("hw", [(1200, 1920), (900, 1600), (1194, 1919), (1034, 1700), (1130, 930), (720, 1280),
                                (1080, 1920), (2160, 3840), (640, 640), (480, 641)])
@pytest.mark.parametrize("auto", [False, True])
def test_preprocess_bit_exact_downscale(cuda_dev, hw, auto):
    f = _frames(hw, hw[0] + hw[1] + auto, B=2)
    ref = olb.preprocess_ref(list(f), (640, 640), auto=auto)
    got = m.preprocess(torch.from_numpy(f).to(cuda_dev), (640, 640), auto=auto)
    assert got.shape == ref.shape
    assert torch.equal(got.cpu(), ref)
    got8 = m.letterbox(torch.from_numpy(f).to(cuda_dev), (640, 640), auto=auto)
    ref8 = np.stack([olb.letterbox_ref(x, (640, 640), auto=auto) for x in f])
    assert np.array_equal(got8.cpu().numpy(), ref8)


def test_golden_frames_sha256(cuda_dev, golden_dir):
    """Real dataset frames (JPEG fixtures) -> same bytes as the real cv2 letterbox recorded in the dev container."""
    gold = json.load(open(os.path.join(golden_dir, "letterbox_golden.json")))
    for key, g in gold.items():
        name, auto = key.split("|auto=")
        im = cv2.imread(os.path.join(golden_dir, "frames", name))
        got = m.letterbox(torch.from_numpy(im).to(cuda_dev), (640, 640), auto=bool(int(auto))).cpu().numpy()
        assert list(got.shape) == g["shape"]
        assert hashlib.sha256(got.tobytes()).hexdigest() == g["sha256"], key


def test_pitched_and_single_image_inputs(cuda_dev):
    f = _frames((300, 500), 7)[0]
    big = np.zeros((300, 512, 3), np.uint8)
    big[:, :500] = f
    view = torch.from_numpy(big).to(cuda_dev)[:, :500]            # row pitch 1536 B, not contiguous
    ref = olb.preprocess_ref([f], (320, 320))
    assert torch.equal(m.preprocess(view, (320, 320)).cpu(), ref)
    assert np.array_equal(m.letterbox(view, (320, 320)).cpu().numpy(), olb.letterbox_ref(f, (320, 320)))
    # no-resize case (already the target size) and non-default padding value
    g = _frames((640, 640), 9)[0]
    assert np.array_equal(m.letterbox(torch.from_numpy(g).to(cuda_dev), (640, 640)).cpu().numpy(), g)
    h = _frames((100, 200), 11)[0]
    assert np.array_equal(m.letterbox(torch.from_numpy(h).to(cuda_dev), (128, 256), padding_value=0).cpu().numpy(),
                          olb.letterbox_ref(h, (128, 256), padding_value=0))


@pytest.mark.parametrize("hw,new", [((543, 770), 1280), ((300, 400), 640), ((37, 53), 640), ((100, 200), (128, 256))])
def test_upscale_bit_exact(cuda_dev, hw, new):
    """Up-scaling (pipe.py's imgsz=1280 on a 770x543 capture): the vertical axis keeps the fractional
    weights of clamped taps (cv::resize clamps row indices only) -- restated, so also bit-exact."""
    f = _frames(hw, 3, B=2)
    new = (new, new) if isinstance(new, int) else new
    got = m.letterbox(torch.from_numpy(f).to(cuda_dev), new, auto=True).cpu().numpy()
    ref = np.stack([olb.letterbox_ref(x, new, auto=True) for x in f])
    assert np.array_equal(got, ref)
    assert torch.equal(m.preprocess(torch.from_numpy(f).to(cuda_dev), new, auto=True).cpu(),
                       olb.preprocess_ref(list(f), new, auto=True))


def test_letterbox_flags_scaleup_center_scalefill(cuda_dev):
    """The remaining LetterBox.__call__ options (scaleup=False, center=False, scale_fill=True, stride 64)."""
    def ul(img, new_shape, auto=False, scale_fill=False, scaleup=True, center=True, stride=32):
        g = olb.letterbox_geometry(img.shape[:2], new_shape, auto, scale_fill, scaleup, center, stride)
        out = img
        if (img.shape[1], img.shape[0]) != (g["new_w"], g["new_h"]):
            out = cv2.resize(img, (g["new_w"], g["new_h"]), interpolation=cv2.INTER_LINEAR)
        return cv2.copyMakeBorder(out, g["top"], g["bottom"], g["left"], g["right"], cv2.BORDER_CONSTANT, value=(114,) * 3)
    for hw, kw in [((300, 400), dict(scaleup=False)), ((900, 1600), dict(center=False)),
                   ((900, 1600), dict(scale_fill=True)), ((1130, 930), dict(auto=True, stride=64)),
                   ((250, 333), dict(scaleup=False, center=False, auto=True))]:
        f = _frames(hw, 5)[0]
        got = m.letterbox(torch.from_numpy(f).to(cuda_dev), (640, 640), **kw).cpu().numpy()
        assert np.array_equal(got, ul(f, (640, 640), **kw)), (hw, kw)


def test_half_output_matches_torch_half_semantics(cuda_dev):
    """half=True form (predict(half=True): `im.half(); im /= 255` on the device): fp16 output equal to torch's own
    fp16 division of the fp16 image by 255 (computed in fp32, rounded once to fp16) on the oracle's uint8 letterbox,
    for a down-scale, an odd shape and a slice-sized identity."""
    from oracle import letterbox as olb
    rng = np.random.default_rng(5)
    for (H, W), imgsz in [((1200, 1920), 640), ((543, 770), 1280), ((640, 640), 640), ((333, 517), 320)]:
        frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
        got = m.preprocess(torch.from_numpy(frames).to(cuda_dev), (imgsz, imgsz), half=True)
        assert got.dtype == torch.float16
        lb = np.stack([olb.letterbox_ref(f, (imgsz, imgsz)) for f in frames])                  # (B,h,w,3) uint8 BGR
        im = torch.from_numpy(np.ascontiguousarray(lb[..., ::-1].transpose(0, 3, 1, 2))).to(cuda_dev).half()
        im /= 255                                                                               # torch's fp16 division on the device
        assert torch.equal(got, im), (H, W, imgsz)
        assert torch.equal(got.cpu(), olb.preprocess_ref(list(frames), (imgsz, imgsz)).half())


def test_round2_golden_frames_test2_and_21_dataset_frames(cuda_dev, golden_dir):
    """BASELINE configs[0]'s literal input (test2.png, read as yolo.py:360 reads it) plus ten dataset frames per
    production size and the 1700x1034 one, both letterbox modes: the uint8 letterbox and the fp32 network input are
    the bytes the real cv2 leaves produced in the dev container (tests/golden/make_golden_r2.py)."""
    gold = json.load(open(os.path.join(golden_dir, "letterbox_golden_r2.json")))
    assert len(gold) == 44
    for key, g in sorted(gold.items()):
        name, auto = key.split("|auto=")
        im = cv2.imread(os.path.join(golden_dir, name))
        assert list(im.shape[:2]) == g["src_hw"]
        d = torch.from_numpy(im).to(cuda_dev)
        got = m.letterbox(d, (640, 640), auto=bool(int(auto))).cpu().numpy()
        assert list(got.shape) == g["shape"]
        assert hashlib.sha256(got.tobytes()).hexdigest() == g["sha256"], key
        net = m.preprocess(d, (640, 640), auto=bool(int(auto))).cpu().numpy()
        assert list(net.shape) == g["net_shape"]
        assert hashlib.sha256(net.tobytes()).hexdigest() == g["net_sha256"], key
