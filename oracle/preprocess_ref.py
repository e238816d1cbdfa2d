"""ORACLE (test infrastructure, NOT product code) -- the input pipeline in front of `tokenize`.

CPU restatement (numpy, fp32) of reference inference/utils.py:12-16 `NPZParser.preprocess`:
    images = images / 255 ; images = torchvision.transforms.functional.resize(images, [S, S])
For float tensors torchvision's resize is `torch.nn.functional.interpolate(mode='bilinear', antialias=True,
align_corners=False)`, i.e. ATen's separable anti-aliased bilinear filter (aten/src/ATen/native/cpu/UpSampleKernel.cpp,
`_compute_indices_min_size_weights_aa` with the triangle filter): per output index i along one axis
    scale   = in / out ;  support = scale if scale >= 1 else 1 ;  center = scale * (i + 0.5)
    xmin    = max(int(center - support + 0.5), 0) ;  xsize = min(int(center + support + 0.5), in) - xmin
    w_j     = tri((j + xmin - center + 0.5) / max(scale, 1)),  tri(x) = max(0, 1 - |x|),  normalised to sum 1
applied first along W, then along H.  (Same arithmetic in ivideogpt/data/simple_dataloader.py:394,510.)

Pinned: tests/golden/preprocess_fractal.npz holds the output of the REFERENCE's own NPZParser.preprocess on frames of its
own fixture inference/samples/fractal_sample.npz (tests/golden/make_golden_preprocess.py); tests/test_preprocess.py checks
this restatement against it on CPU and the CUDA kernel against both on the GPU.
"""
from __future__ import annotations

import numpy as np


def aa_weights(in_size: int, out_size: int):
    """Per output index: (xmin, weights fp32 [xsize])."""
    scale = np.float32(in_size) / np.float32(out_size)
    support = scale if scale >= 1.0 else np.float32(1.0)
    inv = np.float32(1.0) / scale if scale >= 1.0 else np.float32(1.0)
    out = []
    for i in range(out_size):
        center = scale * np.float32(i + 0.5)
        xmin = max(int(np.float32(center - support + np.float32(0.5))), 0)
        xsize = min(int(np.float32(center + support + np.float32(0.5))), in_size) - xmin
        j = np.arange(xsize, dtype=np.float32)
        x = np.abs((j + np.float32(xmin) - center + np.float32(0.5)) * inv).astype(np.float32)
        w = np.where(x < 1.0, np.float32(1.0) - x, np.float32(0.0)).astype(np.float32)
        total = np.float32(0.0)
        for v in w:
            total = np.float32(total + v)
        out.append((xmin, (w / total).astype(np.float32)))
    return out


def preprocess_ref(frames_u8: np.ndarray, size_hw) -> np.ndarray:
    """frames [T, H, W, C] (any integer / float dtype, values 0..255) -> float32 [T, C, S_h, S_w] in [0, 1]."""
    T, H, W, C = frames_u8.shape
    oh, ow = size_hw
    x = (frames_u8.astype(np.float32) / np.float32(255.0)).transpose(0, 3, 1, 2)      # utils.py:13 + :37 permute
    wx, wy = aa_weights(W, ow), aa_weights(H, oh)
    tmp = np.zeros((T, C, H, ow), dtype=np.float32)
    for o, (x0, w) in enumerate(wx):                       # horizontal pass
        acc = x[..., x0] * w[0]
        for j in range(1, len(w)):
            acc = (acc + x[..., x0 + j] * w[j]).astype(np.float32)
        tmp[..., o] = acc
    out = np.zeros((T, C, oh, ow), dtype=np.float32)
    for o, (y0, w) in enumerate(wy):                       # vertical pass
        acc = tmp[:, :, y0, :] * w[0]
        for j in range(1, len(w)):
            acc = (acc + tmp[:, :, y0 + j, :] * w[j]).astype(np.float32)
        out[:, :, o, :] = acc
    return out
