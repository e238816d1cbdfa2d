"""Regenerates tests/golden/*.npz from the CPU oracle (seeded weights, seeded inputs).  Run from the repo root:
    python tests/golden/make_golden.py
The reference itself cannot produce these vectors offline (diffusers is not installable here)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.vq_model_ref import TINY_CFG, RefCompressiveVQModel, seeded_init_  # noqa: E402

torch.set_num_threads(1)   # summation order of the CPU kernels must not depend on the thread count
ref = seeded_init_(RefCompressiveVQModel(**TINY_CFG).eval())
px = torch.rand(1, 5, 3, 64, 64, generator=torch.Generator().manual_seed(11))
tok, lab = ref.tokenize(px, 2)
rec = ref.detokenize(tok, 2)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tokenizer_tiny.npz"), pixels=px.numpy(),
                    tokens=tok.numpy(), labels=lab.numpy(), recon=rec.numpy().astype(np.float32))
print("tokens", tok.shape, "recon", rec.shape, float(rec.mean()))
