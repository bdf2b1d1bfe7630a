"""Video front-end of the SG-RLVR step: the integer logic of the reference's vendored qwen-vl-utils
(/root/reference/SpaceR-SG-RLVR/src/qwen-vl-utils/src/qwen_vl_utils/vision_process.py, "QVU") and the HF video
processor's rescale/normalise/patchify on the GPU (sb_video_patchify).

Host side (pure Python integers, checked against the reference's own functions via tests/golden/vision.json):
    smart_resize      QVU:61-87     target (h, w): multiples of 28 inside [min_pixels, max_pixels]
    smart_nframes     QVU:145-182   frame count from fps / min / max, floored to a multiple of 2
    frame_indices     QVU:246       torch.linspace(0, total - 1, nframes).round().long()
Device side:
    patchify(frames[, perm])  ->  pixel_values_videos (bf16 for the ViT and/or fp32 bit-exact with HF), video_grid_thw
    resize_frames(frames, h, w)  ->  the bicubic-antialias resize of fetch_video (QVU:310-315) on the GPU
Container decode (QVU:185-256, `_read_video_decord` / `_read_video_torchvision`): `read_video` opens the file with
OpenCV's FFmpeg backend -- decord and PyAV are not in this image -- and decodes ONLY the sampled frames; `fetch_video`
chains it with the GPU resize, i.e. the whole of the reference's `fetch_video` for a video path.  (Decoding on NVDEC is
not built: SURVEY 8(f) row 2 remainder.)
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import ops
from .ops import SpacerError

IMAGE_FACTOR = 28
MIN_PIXELS = 4 * 28 * 28
MAX_PIXELS = 256 * 28 * 28
MAX_RATIO = 200
VIDEO_MIN_PIXELS = 128 * 28 * 28        # QVU:32-33 (the SpaceR fork pins min == max == 128*28*28)
VIDEO_MAX_PIXELS = 128 * 28 * 28
VIDEO_TOTAL_PIXELS = int(128000 * 28 * 28 * 0.9)   # QVU:42 (default of the VIDEO_MAX_PIXELS environment variable)
FRAME_FACTOR = 2
FPS = 2.0
FPS_MIN_FRAMES = 4
FPS_MAX_FRAMES = 16
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)      # transformers/utils/constants.py:5-6
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def round_by_factor(number, factor):
    return round(number / factor) * factor           # Python round: half to even, like the reference (QVU:48)


def ceil_by_factor(number, factor):
    return math.ceil(number / factor) * factor


def floor_by_factor(number, factor):
    return math.floor(number / factor) * factor


def smart_resize(height, width, factor=IMAGE_FACTOR, min_pixels=MIN_PIXELS, max_pixels=MAX_PIXELS):
    """QVU:61-87."""
    if max(height, width) / min(height, width) > MAX_RATIO:
        raise ValueError(f"absolute aspect ratio must be smaller than {MAX_RATIO}, got {max(height, width) / min(height, width)}")
    h_bar = max(factor, round_by_factor(height, factor))
    w_bar = max(factor, round_by_factor(width, factor))
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = floor_by_factor(height / beta, factor)
        w_bar = floor_by_factor(width / beta, factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = ceil_by_factor(height * beta, factor)
        w_bar = ceil_by_factor(width * beta, factor)
    return h_bar, w_bar


def smart_nframes(ele: dict, total_frames: int, video_fps) -> int:
    """QVU:145-182."""
    if "fps" in ele and "nframes" in ele:
        raise AssertionError("Only accept either `fps` or `nframes`")
    if "nframes" in ele:
        nframes = round_by_factor(ele["nframes"], FRAME_FACTOR)
    else:
        fps = ele.get("fps", FPS)
        min_frames = ceil_by_factor(ele.get("min_frames", FPS_MIN_FRAMES), FRAME_FACTOR)
        max_frames = floor_by_factor(ele.get("max_frames", min(FPS_MAX_FRAMES, total_frames)), FRAME_FACTOR)
        nframes = total_frames / video_fps * fps
        nframes = min(min(max(nframes, min_frames), max_frames), total_frames)
        nframes = floor_by_factor(nframes, FRAME_FACTOR)
    if not (FRAME_FACTOR <= nframes and nframes <= total_frames):
        raise ValueError(f"nframes should in interval [{FRAME_FACTOR}, {total_frames}], but got {nframes}.")
    return nframes


def frame_indices(total_frames: int, nframes: int) -> list[int]:
    """QVU:246 (decord path) / QVU:213 (torchvision path)."""
    return torch.linspace(0, total_frames - 1, nframes).round().long().tolist()


def fused_mean_std(rescale_factor: float = 1 / 255, mean=CLIP_MEAN, std=CLIP_STD):
    """HF's fused rescale+normalise constants (image_processing_backends.py:292-306): fp32 mean * (1 / rescale_factor)."""
    m = torch.tensor(mean) * (1.0 / rescale_factor)
    s = torch.tensor(std) * (1.0 / rescale_factor)
    return m.tolist(), s.tolist()


def patchify(frames: torch.Tensor, perm: torch.Tensor | None = None, *, patch=14, t_patch=2, merge=2, want_f32=False,
             want_bf16=True, rescale_factor=1 / 255, mean=CLIP_MEAN, std=CLIP_STD):
    """frames: CUDA uint8 or float32 [F, 3, H, W] with H, W multiples of patch*merge.  perm: optional int32 [F] on the
    device.  Returns (pixel_values bf16 or None, pixel_values fp32 or None, video_grid_thw LongTensor [1, 3])."""
    if not frames.is_cuda or frames.dtype not in (torch.uint8, torch.float32) or frames.dim() != 4:
        raise SpacerError("patchify: frames must be a CUDA uint8/float32 tensor [F, C, H, W]")
    frames = frames.contiguous()
    F_, Cc, H, W = frames.shape
    gt, gh, gw = -(-F_ // t_patch), H // patch, W // patch
    n, row = gt * gh * gw, Cc * t_patch * patch * patch
    o32 = torch.empty((n, row), device=frames.device, dtype=torch.float32) if want_f32 else None
    o16 = torch.empty((n, row), device=frames.device, dtype=torch.bfloat16) if want_bf16 else None
    m, s = fused_mean_std(rescale_factor, mean, std)
    mh = (C.c_float * 4)(*(m + [0.0] * (4 - len(m))))
    sh = (C.c_float * 4)(*(s + [1.0] * (4 - len(s))))
    if perm is not None and (perm.dtype != torch.int32 or not perm.is_cuda or perm.numel() != F_):
        raise SpacerError("patchify: perm must be a CUDA int32 tensor with one entry per frame")
    ops.call("sb_video_patchify", frames, int(frames.dtype == torch.uint8), F_, Cc, H, W, perm,
             C.cast(mh, C.c_void_p), C.cast(sh, C.c_void_p), patch, t_patch, merge, o32, o16)
    return o16, o32, torch.tensor([[gt, gh, gw]])


# ------------------------------------------------------------------------------------------------------------------
# bicubic antialiased resize (QVU:310-315 -> torchvision resize -> ATen _upsample_bicubic2d_aa)
# ------------------------------------------------------------------------------------------------------------------
_AA_CACHE: dict = {}


def aa_weight_table(in_size: int, out_size: int):
    """Per output index: first source index, tap count and normalised cubic weights (a = -0.5), computed like ATen's
    HelperInterpBase::_compute_indices_min_size_weights_aa for float32 tensors: scale, support, centre and the weights
    in fp32, the filter argument in double.  Returns (weights float32 [out, taps], xmin int32 [out], xsize int32 [out])."""
    import numpy as np
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    support = f32(2.0) * scale if scale >= 1.0 else f32(2.0)
    invscale = f32(1.0) / scale if scale >= 1.0 else f32(1.0)
    taps = int(math.ceil(float(support))) * 2 + 1
    w = np.zeros((out_size, taps), dtype=np.float32)
    xmin = np.zeros(out_size, dtype=np.int32)
    xsize = np.zeros(out_size, dtype=np.int32)
    a = f32(-0.5)

    def filt(x):
        x = f32(abs(x))
        if x < 1.0:
            return ((a + f32(2.0)) * x - (a + f32(3.0))) * x * x + f32(1.0)
        if x < 2.0:
            return (((x - f32(5.0)) * x + f32(8.0)) * x - f32(4.0)) * a
        return f32(0.0)

    for i in range(out_size):
        center = scale * f32(i + 0.5)
        lo = max(int(center - support + f32(0.5)), 0)
        n = min(int(center + support + f32(0.5)), in_size) - lo
        total = f32(0.0)
        for j in range(n):
            arg = (float(f32(j + lo) - center) + 0.5) * float(invscale)       # promoted to double by the 0.5 literal
            w[i, j] = filt(f32(arg))
            total = f32(total + w[i, j])
        if total != 0.0:
            w[i, :n] = w[i, :n] / total
        xmin[i], xsize[i] = lo, n
    return w, xmin, xsize


def resize_frames(frames: torch.Tensor, height: int, width: int, *, out_u8: bool = False, use_fma: bool = True):
    """frames: CUDA uint8 or float32 [F, C, H, W] -> [F, C, height, width], float32 holding integers in 0..255 (what
    fetch_video returns) or uint8.  Bicubic, antialias, round-half-even + clamp as torchvision does for uint8 input."""
    if not frames.is_cuda or frames.dtype not in (torch.uint8, torch.float32) or frames.dim() != 4:
        raise SpacerError("resize_frames: frames must be a CUDA uint8/float32 tensor [F, C, H, W]")
    frames = frames.contiguous()
    F_, Cc, H, W = frames.shape
    dev = frames.device
    tabs = []
    for in_size, out_size in ((W, width), (H, height)):
        key = (in_size, out_size, str(dev))
        if key not in _AA_CACHE:
            w, lo, n = aa_weight_table(in_size, out_size)
            _AA_CACHE[key] = (torch.from_numpy(w).to(dev), torch.from_numpy(lo).to(dev), torch.from_numpy(n).to(dev), w.shape[1])
        tabs.append(_AA_CACHE[key])
    (wh, xh, nh, th), (wv, xv, nv, tv) = tabs
    tmp = torch.empty((F_ * Cc, H, width), device=dev, dtype=torch.float32)
    out = torch.empty((F_, Cc, height, width), device=dev, dtype=torch.uint8 if out_u8 else torch.float32)
    round_u8 = int(frames.dtype == torch.uint8 or out_u8)
    ops.call("sb_resize_bicubic_aa", frames, int(frames.dtype == torch.uint8), F_ * Cc, H, W, tmp, out, int(out_u8), height,
             width, wh, xh, nh, th, wv, xv, nv, tv, round_u8, int(use_fma))
    return out


def video_target_size(nframes: int, height: int, width: int, ele: dict | None = None, image_factor: int = IMAGE_FACTOR):
    """(resized_height, resized_width) exactly as fetch_video computes them (QVU:288-309): max_pixels from the per-video
    pixel budget, then smart_resize; explicit `resized_height` / `resized_width` in `ele` win."""
    ele = ele or {}
    min_pixels = ele.get("min_pixels", VIDEO_MIN_PIXELS)
    total_pixels = ele.get("total_pixels", VIDEO_TOTAL_PIXELS)
    max_pixels = max(min(VIDEO_MAX_PIXELS, total_pixels / nframes * FRAME_FACTOR), int(min_pixels * 1.05))
    max_pixels = min(ele.get("max_pixels", max_pixels), max_pixels)
    if "resized_height" in ele and "resized_width" in ele:
        return smart_resize(ele["resized_height"], ele["resized_width"], factor=image_factor)
    return smart_resize(height, width, factor=image_factor, min_pixels=min_pixels, max_pixels=max_pixels)


def fetch_video_frames(video: torch.Tensor, video_fps: float, ele: dict | None = None):
    """The part of qwen-vl-utils' fetch_video after the container is open (QVU:228-256, 279-318), on the GPU: pick
    `smart_nframes` frames at linspace indices out of the decoded clip `video` (CUDA uint8 [total_frames, C, H, W]),
    resize them (bicubic, antialias) to the `smart_resize` target and return (float32 frames [nframes, C, h, w] holding
    0..255 like the reference, sample_fps).  Decoding the container itself (decord / NVDEC) is not part of this."""
    ele = ele or {}
    total = video.shape[0]
    n = smart_nframes(ele, total_frames=total, video_fps=video_fps)
    idx = torch.tensor(frame_indices(total, n), device=video.device)
    frames = video.index_select(0, idx)
    h, w = video_target_size(n, video.shape[2], video.shape[3], ele)
    sample_fps = n / max(total, 1e-6) * video_fps
    return resize_frames(frames, h, w), sample_fps


# ------------------------------------------------------------------------------------------------------------------
# container decode (QVU:185-256) + the reference's fetch_video for a path (QVU:279-318)
# ------------------------------------------------------------------------------------------------------------------
def read_video(ele: dict):
    """`_read_video_decord` / `_read_video_torchvision` (QVU:185-256): open `ele["video"]` (a path, "file://" allowed),
    pick `smart_nframes` frames at `linspace(0, total - 1, n).round()` and return (uint8 CPU tensor [n, 3, H, W] RGB,
    sample_fps).  `video_start` / `video_end` (seconds) restrict the clip like the torchvision backend does.  Only the
    selected frames are colour-converted and kept; the frames between them are skipped with `grab()` (no decode output)."""
    import cv2
    path = ele["video"]
    if not isinstance(path, str):
        raise SpacerError("read_video: ele['video'] must be a path")
    if path.startswith("file://"):
        path = path[7:]
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise SpacerError(f"read_video: cannot open {path}")
    try:
        n_total = int(round(cap.get(cv2.CAP_PROP_FRAME_COUNT)))
        video_fps = float(cap.get(cv2.CAP_PROP_FPS)) or FPS
        first = 0
        last = n_total - 1
        if ele.get("video_start") is not None:
            first = min(max(int(math.ceil(float(ele["video_start"]) * video_fps)), 0), max(n_total - 1, 0))
        if ele.get("video_end") is not None:
            last = min(int(math.floor(float(ele["video_end"]) * video_fps)), n_total - 1)
        total = last - first + 1
        if n_total <= 0 or total <= 0:
            raise SpacerError(f"read_video: {path} holds no frames in the requested range")
        n = smart_nframes(ele, total_frames=total, video_fps=video_fps)
        wanted = [first + i for i in frame_indices(total, n)]
        frames, pos, k = [], 0, 0
        if wanted[0] > 64:                       # long lead-in: seek instead of grabbing frame by frame
            cap.set(cv2.CAP_PROP_POS_FRAMES, wanted[0])
            pos = wanted[0]
        while k < len(wanted):
            target = wanted[k]
            while pos < target:                  # skip without retrieving
                if not cap.grab():
                    raise SpacerError(f"read_video: {path} ended at frame {pos} (container reports {n_total})")
                pos += 1
            ok, bgr = cap.read()
            if not ok:
                raise SpacerError(f"read_video: {path} ended at frame {pos} (container reports {n_total})")
            pos += 1
            rgb = torch.from_numpy(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)).permute(2, 0, 1).contiguous()
            frames.append(rgb)
            k += 1
            while k < len(wanted) and wanted[k] == target:      # linspace can repeat an index on very short clips
                frames.append(rgb)
                k += 1
    finally:
        cap.release()
    sample_fps = n / max(total, 1e-6) * video_fps
    return torch.stack(frames), sample_fps


def fetch_video(ele: dict, device="cuda", image_factor: int = IMAGE_FACTOR, return_video_sample_fps: bool = False):
    """qwen-vl-utils' `fetch_video` for a video path (QVU:279-318): decode + sample (read_video), then the pixel budget /
    `smart_resize` target and the bicubic-antialias resize on the GPU.  Returns float32 frames [n, 3, h, w] holding 0..255
    (what the reference hands to the HF processor), optionally with sample_fps."""
    video, sample_fps = read_video(ele)
    video = video.pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else video.to(device)
    n, _, height, width = video.shape
    h, w = video_target_size(n, height, width, ele, image_factor)
    out = resize_frames(video, h, w)
    return (out, sample_fps) if return_video_sample_fps else out
