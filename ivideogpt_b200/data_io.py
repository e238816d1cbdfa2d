"""Input pipeline in front of `tokenize` -- the `NPZParser` surface of reference inference/utils.py:6-41 with the
arithmetic (`/ 255` + antialiased bilinear resize, utils.py:12-16) on the GPU (csrc/elementwise.cu::resize_aa_kernel).

    parser = NPZParser(segment_length=16, image_size=64)
    frames, actions = parser.parse("episode.npz", "fractal20220817_data", load_action=False)   # CUDA fp32 [T,3,S,S] in [0,1]

Differences to the reference, on purpose: the episode's uint8 frames are copied to the device as uint8 (4x fewer bytes than
the reference's float tensor) and converted / resized there; `preprocess` also accepts the reference's own argument (a float
[T,C,H,W] tensor in 0..255) as long as it is a CUDA tensor.  The per-dataset tables of the reference (BASE_STEPSIZE,
DISPLAY_KEY, utils.py:44-86) are data, not arithmetic: pass them in (`from inference.utils import BASE_STEPSIZE, DISPLAY_KEY`)
or rely on the defaults (step size 1, key 'image').  SURVEY.md section 8(f) rank 4.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import ops


class NPZParser:
    def __init__(self, segment_length, image_size=64, base_stepsize: Optional[Dict[str, int]] = None,
                 display_key: Optional[Dict[str, str]] = None, device="cuda"):
        self.segment_length = segment_length
        self.image_size = image_size
        self.base_stepsize = dict(base_stepsize or {})
        self.display_key = dict(display_key or {})
        self.device = torch.device(device)

    def preprocess(self, images: torch.Tensor) -> torch.Tensor:
        """utils.py:12-16.  images: CUDA uint8 [T,H,W,C] (episode layout) or CUDA float [T,C,H,W] in 0..255 (the reference's
        argument) -> fp32 [T,C,S,S] in [0,1]."""
        size = [self.image_size, self.image_size]
        if images.dtype == torch.uint8:
            return ops.preprocess_resize(images, size, channels_last=True)
        return ops.preprocess_resize(images.float(), size, channels_last=False)

    def get_segment(self, episode, actions, stepsize=1):
        """utils.py:18-27: random window of segment_length frames with the given stride (shrunk for short episodes)."""
        n = len(episode)
        if stepsize * self.segment_length > n:
            stepsize = max(1, n // self.segment_length)
        span = stepsize * self.segment_length
        start = np.random.randint(max(n - span + 1, 1))
        sel = slice(start, start + span, stepsize)
        return episode[sel], (actions[sel] if actions is not None else None)

    def get_stepsize(self, dataset_name):
        """utils.py:29-30: frame stride relative to the fractal dataset's."""
        base = self.base_stepsize
        if not base:
            return 1
        return max(round(base.get(dataset_name, 1) / base.get("fractal20220817_data", 1)), 1)

    def parse(self, npz_file, dataset_name, load_action=False):
        """utils.py:32-41 -> (frames fp32 CUDA [T,3,S,S] in [0,1], actions fp32 CUDA [T,A] or None)."""
        with np.load(npz_file) as z:
            images = z[self.display_key.get(dataset_name, "image")]
            actions = z["action"] if load_action else None
        images, actions = self.get_segment(images, actions, self.get_stepsize(dataset_name))
        u8 = np.ascontiguousarray(images).astype(np.uint8, copy=False)          # episodes store 0..255 (some as int64)
        frames = torch.from_numpy(u8).pin_memory().to(self.device, non_blocking=True)
        out = self.preprocess(frames)
        act = torch.as_tensor(np.asarray(actions), dtype=torch.float32).to(self.device) if actions is not None else None
        return out, act
