"""The prefill's fused attention launch at the cfg64 shape (B = 64 clips, 12 heads, 514 prompt tokens, head_dim 64, bf16) for
`ncu --set full -k regex:flash_attn_kernel`, plus a CUDA-event timing of the same launch (printed; never taken under ncu)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ivideogpt_b200 import ops

dev = torch.device("cuda:0")
B, H, L, Lmax = 64, 12, 514, 752
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(B, H, L, 64, device=dev, generator=g).to(torch.bfloat16)
k = torch.randn(B, H, Lmax, 64, device=dev, generator=g).to(torch.bfloat16)
vt = torch.randn(B, H, 64, Lmax, device=dev, generator=g).to(torch.bfloat16)
out = torch.zeros(B * L, H * 64, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.flash_attn(q, k, vt, out, B, H, L, L, Lmax * 64, 64 * Lmax, Lmax)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.flash_attn(q, k, vt, out, B, H, L, L, Lmax * 64, 64 * Lmax, Lmax)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flops = 4.0 * B * H * 64 * (L * (L + 1) / 2)            # causal: QK^T and PV over the lower triangle
print(json.dumps({"shape": [B, H, L, 64], "ms": ms, "causal_TFLOPs": flops / ms / 1e9,
                  "algorithmic_bytes": 2.0 * B * H * L * 64 * 4, "GBps_on_algorithmic_bytes": 2.0 * B * H * L * 64 * 4 / ms / 1e6}))
