"""Regenerates tests/golden/preprocess_fractal.npz with the REFERENCE's own inference/utils.py NPZParser.preprocess
(imported by file path; needs torch + torchvision, both in this image) on frames of the reference's own fixture
inference/samples/fractal_sample.npz (uint8 [22, 256, 320, 3]).  Run from the repo root in the build container:
    python tests/golden/make_golden_preprocess.py"""
import importlib.util
import os

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/inference/utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

torch.set_num_threads(1)
frames = np.load("/root/reference/inference/samples/fractal_sample.npz")["image"][[0, 7, 21]]      # 3 of 22 frames
x = torch.Tensor(np.array(frames)).permute(0, 3, 1, 2)                  # utils.py:37
out64 = ref.NPZParser(16, 64).preprocess(x).numpy()
out256 = ref.NPZParser(16, 256).preprocess(x).numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess_fractal.npz"), frames=frames,
                    out64=out64.astype(np.float32), out256=out256[:1].astype(np.float32),
                    versions=np.array([torch.__version__, torchvision.__version__]))
print(frames.shape, out64.shape, out256.shape, float(out64.mean()))
