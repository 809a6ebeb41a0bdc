"""CPU restatement of the per-frame host glue around the network forward.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the product
path (mmt_b200/frames.py -> csrc/frames.cu) never calls it.

What is restated, with the reference lines it follows:
  * crop geometry, zero padding and the "drop the last row/column" quirk of `sample_target`
    (lib/train/data/processing_utils.py:15-83);
  * the resize inside it, `cv.resize(im, (S, S))` on uint8 = OpenCV's fixed-point INTER_LINEAR.  OpenCV is a
    third-party dependency of the reference (opencv-python, unpinned in install_pytorch17.sh; 4.13.0 in this image);
    the arithmetic below restates its published algorithm (modules/imgproc/src/resize.cpp: 11-bit coefficient
    tables, HResizeLinear / VResizeLinear<uchar>) and is pinned bit for bit against cv2 itself in
    tests/test_frames_oracle.py and against fixtures produced by the reference's own functions
    (oracle/gen_golden_frames.py -> tests/golden/frames_*.npz);
  * `Preprocessor_Multimodal.process` (lib/test/tracker/tracker_utils.py:37-48): JET colour map on the infrared crop
    (cv2.applyColorMap: BGR2GRAY fixed point + 256-entry table), /255, mean/std, HWC -> CHW;
  * the box update of `track()` (lib/test/tracker/asymmetric_shared_ce.py:99-103,134-140): prediction scaled in
    fp32, mapped back in float64, `clip_box` (lib/utils/box_ops.py:155-164) with margin 10.
"""
from __future__ import annotations

import math

import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)

COEF_BITS = 11                      # INTER_RESIZE_COEF_BITS
COEF_ONE = 1 << COEF_BITS           # INTER_RESIZE_COEF_SCALE


# cv2.applyColorMap(np.arange(256, dtype=np.uint8)[None], cv2.COLORMAP_JET)[0].tobytes().hex() - OpenCV 4.13.0 (its
# colormap.cpp holds the Jet map as 256 literal float triples; this is that table after convertTo(CV_8U, 255)).
# Data, not an algorithm: tests/test_frames_oracle.py checks it against the live cv2 and against the product's copy.
_JET_HEX = (
    "8000008400008800008c00009000009400009800009c0000a00000a40000a80000ac0000b00000b40000b80000bc0000"
    "c00000c40000c80000cc0000d00000d40000d80000dc0000e00000e40000e80000ec0000f00000f40000f80000fc0000"
    "ff0000ff0400ff0800ff0c00ff1000ff1400ff1800ff1c00ff2000ff2400ff2800ff2c00ff3000ff3400ff3800ff3c00"
    "ff4000ff4400ff4800ff4c00ff5000ff5400ff5800ff5c00ff6000ff6400ff6800ff6c00ff7000ff7400ff7800ff7c00"
    "ff8000ff8400ff8800ff8c00ff9000ff9400ff9800ff9c00ffa000ffa400ffa800ffac00ffb000ffb400ffb800ffbc00"
    "ffc000ffc400ffc800ffcc00ffd000ffd400ffd800ffdc00ffe000ffe400ffe800ffec00fff000fff400fff800fffc00"
    "feff02faff06f6ff0af2ff0eeeff12eaff16e6ff1ae2ff1edeff22daff26d6ff2ad2ff2eceff32caff36c6ff3ac2ff3e"
    "beff42baff46b6ff4ab2ff4eaeff52aaff56a6ff5aa2ff5e9eff629aff6696ff6a92ff6e8eff728aff7686ff7a82ff7e"
    "7eff827aff8676ff8a72ff8e6eff926aff9666ff9a62ff9e5effa25affa656ffaa52ffae4effb24affb646ffba42ffbe"
    "3effc23affc636ffca32ffce2effd22affd626ffda22ffde1effe21affe616ffea12ffee0efff20afff606fffa01fffe"
    "00fcff00f8ff00f4ff00f0ff00ecff00e8ff00e4ff00e0ff00dcff00d8ff00d4ff00d0ff00ccff00c8ff00c4ff00c0ff"
    "00bcff00b8ff00b4ff00b0ff00acff00a8ff00a4ff00a0ff009cff0098ff0094ff0090ff008cff0088ff0084ff0080ff"
    "007cff0078ff0074ff0070ff006cff0068ff0064ff0060ff005cff0058ff0054ff0050ff004cff0048ff0044ff0040ff"
    "003cff0038ff0034ff0030ff002cff0028ff0024ff0020ff001cff0018ff0014ff0010ff000cff0008ff0004ff0000ff"
    "0000fc0000f80000f40000f00000ec0000e80000e40000e00000dc0000d80000d40000d00000cc0000c80000c40000c0"
    "0000bc0000b80000b40000b00000ac0000a80000a40000a000009c00009800009400009000008c000088000084000080")


def jet_lut() -> np.ndarray:
    """[256, 3] uint8, channel order as cv2.applyColorMap returns it (B, G, R)."""
    return np.frombuffer(bytes.fromhex(_JET_HEX), dtype=np.uint8).reshape(256, 3).copy()


def crop_geometry(box, factor: float, frame_h: int, frame_w: int):
    """processing_utils.py:31-49 - integer crop window and the valid frame range that survives the slicing.

    Returns crop_sz, x1, y1 and the half-open frame ranges [xa, xb) x [ya, yb) that are copied; everything else of
    the crop_sz x crop_sz window is zero.  (`x2_pad = max(x2 - W + 1, 0)`: when the window reaches the last column
    that column is dropped too - restated as is.)"""
    x, y, w, h = [float(v) for v in box]
    crop_sz = math.ceil(math.sqrt(w * h) * factor)
    if crop_sz < 1:
        raise Exception("Too small bounding box.")
    x1 = int(round(x + 0.5 * w - crop_sz * 0.5))        # Python round: half to even
    y1 = int(round(y + 0.5 * h - crop_sz * 0.5))
    x2, y2 = x1 + crop_sz, y1 + crop_sz
    xa, xb = max(x1, 0), x2 - max(x2 - frame_w + 1, 0)
    ya, yb = max(y1, 0), y2 - max(y2 - frame_h + 1, 0)
    return crop_sz, x1, y1, xa, xb, ya, yb


def padded_crop(im: np.ndarray, box, factor: float):
    """The crop_sz x crop_sz x C uint8 window before the resize (processing_utils.py:51-57)."""
    H, W = im.shape[:2]
    crop_sz, x1, y1, xa, xb, ya, yb = crop_geometry(box, factor, H, W)
    out = np.zeros((crop_sz, crop_sz, im.shape[2]), dtype=np.uint8)
    if xb > xa and yb > ya:
        out[ya - y1:yb - y1, xa - x1:xb - x1] = im[ya:yb, xa:xb]
    return out, crop_sz


def _round_half_even_to_short(v: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(v), -32768, 32767).astype(np.int32)


def linear_tables(src: int, dst: int):
    """resize.cpp (resize setup loop): per destination index the source index and the two 11-bit weights.

    scale = 1 / (dst / src) in double; f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s."""
    inv = float(dst) / float(src)
    scale = 1.0 / inv
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_linear_u8(src: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(src, (dst_w, dst_h)) for uint8 HWC, INTER_LINEAR, bit for bit.

    Horizontal: x < 0 -> (0, weight 0); x + 1 >= width -> single tap `src[min(x, width-1)] * 2048`.
    Vertical: weights kept, row indices clamped.  Accumulation exactly as VResizeLinear<uchar>:
    ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2."""
    sh, sw = src.shape[:2]
    sx, fx = linear_tables(sw, dst_w)
    sy, fy = linear_tables(sh, dst_h)
    # horizontal tables
    neg = sx < 0
    fx = np.where(neg, np.float32(0), fx)
    sx = np.where(neg, 0, sx)
    edge = sx + 1 >= sw
    a0 = _round_half_even_to_short((np.float32(1.0) - fx) * np.float32(COEF_ONE))
    a1 = _round_half_even_to_short(fx * np.float32(COEF_ONE))
    sx0 = np.minimum(sx, sw - 1)
    sx1 = np.minimum(sx + 1, sw - 1)
    a0 = np.where(edge, COEF_ONE, a0)
    a1 = np.where(edge, 0, a1)
    b0 = _round_half_even_to_short((np.float32(1.0) - fy) * np.float32(COEF_ONE))
    b1 = _round_half_even_to_short(fy * np.float32(COEF_ONE))
    r0 = np.clip(sy, 0, sh - 1)
    r1 = np.clip(sy + 1, 0, sh - 1)
    s32 = src.astype(np.int32)
    hrow = s32[:, sx0, :] * a0[None, :, None] + s32[:, sx1, :] * a1[None, :, None]      # [sh, dst_w, C]
    S0 = hrow[r0]
    S1 = hrow[r1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def sample_target(im: np.ndarray, box, factor: float, output_sz: int):
    """processing_utils.py:15-83 without the mask outputs: (uint8 crop [S, S, C], resize_factor)."""
    crop, crop_sz = padded_crop(im, box, factor)
    return resize_linear_u8(crop, output_sz, output_sz), output_sz / crop_sz


def bgr2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for uint8 (color_rgb.simd.hpp: 15-bit coefficients, channel 0 = 'B')."""
    i = img.astype(np.int32)
    return ((i[..., 0] * 3735 + i[..., 1] * 19235 + i[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def apply_jet(img: np.ndarray) -> np.ndarray:
    """cv2.applyColorMap(img, COLORMAP_JET) on a 3-channel uint8 image: gray first, then the table (tracker_utils.py:43)."""
    return jet_lut()[bgr2gray_u8(img)]


def normalize(img_u8: np.ndarray) -> np.ndarray:
    """tracker_utils.py:44-47: ((x / 255) - mean) / std in fp32, HWC -> CHW."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = (x - MEAN) / STD
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def process_multimodal(crop_v: np.ndarray, crop_i: np.ndarray):
    return normalize(crop_v), normalize(apply_jet(crop_i))


def clip_box(box, H, W, margin=0):
    """lib/utils/box_ops.py:155-164."""
    x1, y1, w, h = box
    x2, y2 = x1 + w, y1 + h
    x1 = min(max(0, x1), W - margin)
    x2 = min(max(margin, x2), W)
    y1 = min(max(0, y1), H - margin)
    y2 = min(max(margin, y2), H)
    w = max(margin, x2 - x1)
    h = max(margin, y2 - y1)
    return [x1, y1, w, h]


def update_state(state, pred_box_f32: np.ndarray, resize_factor: float, search_size: int, H: int, W: int, margin=10):
    """lib/test/tracker/asymmetric_shared_ce.py:99-103,134-140.

    `pred_boxes.mean(0) * search_size / resize_factor` is fp32 tensor arithmetic (one box: the mean is the box), then
    `.tolist()` widens to float64 and everything after is Python float arithmetic."""
    p = np.asarray(pred_box_f32, dtype=np.float32)
    p = (p * np.float32(search_size)) / np.float32(resize_factor)
    cx, cy, w, h = [float(v) for v in p]
    cx_prev = state[0] + 0.5 * state[2]
    cy_prev = state[1] + 0.5 * state[3]
    half_side = 0.5 * search_size / resize_factor
    cx_real = cx + (cx_prev - half_side)
    cy_real = cy + (cy_prev - half_side)
    return clip_box([cx_real - 0.5 * w, cy_real - 0.5 * h, w, h], H, W, margin=margin)


def sigmoid_f32(logit) -> float:
    """`pred_scores.view(1).sigmoid().item()`: fp32 sigmoid widened to a Python float."""
    x = np.float32(logit)
    return float(np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32)))


def online_score_step(max_score: float, logit, decay: float = 1.0):
    """lib/test/tracker/mixformer_convmae_online.py:99,105-113: returns (new max_pred_score, take) where `take` means the
    crop at the new box becomes `online_max_template`."""
    s = sigmoid_f32(logit)
    m = max_score * decay
    take = s > 0.5 and s > m
    return (s if take else m), take


class OnlineTrackerOracle:
    """MixFormerOnline with online_size == 1 (mixformer_convmae_online.py:62-128) for one RGB sequence; `network` is a
    callable (template, online_template, search) -> (pred_box_f32[4] cxcywh, logit)."""

    def __init__(self, network, template_factor, template_size, search_factor, search_size, update_interval,
                 max_score_decay=1.0, rgbt=False, online_size=1):
        # rgbt: the RGB-T online class (lib/test/tracker/asymmetric_shared_online.py:62-119): images and crops are [v, i]
        # pairs, Preprocessor_Multimodal (JET on the infrared crop), same bookkeeping
        self.net = network
        self.tf, self.ts, self.sf, self.ss = template_factor, template_size, search_factor, search_size
        self.update_interval, self.decay, self.rgbt = update_interval, max_score_decay, rgbt
        # online_size > 1 (mixformer_convmae_online.py:66-68,94-97,115-124): the online templates form a list that grows to
        # online_size and is then overwritten round-robin; `network` receives the stacked list [n, 3, T, T]
        self.online_size = online_size

    def _crop(self, image, state, factor, size):
        if self.rgbt:
            cv_, rf = sample_target(image[0], state, factor, size)
            ci_, _ = sample_target(image[1], state, factor, size)
            return list(process_multimodal(cv_, ci_)), rf
        c, rf = sample_target(image, state, factor, size)
        return normalize(c), rf

    def initialize(self, image, init_box):
        self.template, _ = self._crop(image, list(init_box), self.tf, self.ts)
        self.online_template = self.template if self.online_size == 1 else np.stack([self.template])
        self.online_forget_id = 0
        self.online_max_template = self.template
        self.max_pred_score = -1.0
        self.state = [float(v) for v in init_box]
        self.frame_id = 0

    def track(self, image):
        H, W = (image[0] if self.rgbt else image).shape[:2]
        self.frame_id += 1
        search, rf = self._crop(image, self.state, self.sf, self.ss)
        pred, logit = self.net(self.template, self.online_template, search)
        self.state = update_state(self.state, pred, rf, self.ss, H, W, margin=10)
        self.max_pred_score, take = online_score_step(self.max_pred_score, logit, self.decay)
        if take:
            self.online_max_template, _ = self._crop(image, self.state, self.tf, self.ts)
        if self.frame_id % self.update_interval == 0:
            if self.online_size == 1:
                self.online_template = self.online_max_template
            elif self.online_template.shape[0] < self.online_size:
                self.online_template = np.concatenate([self.online_template, self.online_max_template[None]])
            else:
                self.online_template = self.online_template.copy()
                self.online_template[self.online_forget_id] = self.online_max_template
                self.online_forget_id = (self.online_forget_id + 1) % self.online_size
            self.max_pred_score = -1
            self.online_max_template = self.template
        return self.state
