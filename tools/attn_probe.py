"""How fast is the standalone fused decode-attention kernel (normal launch, many CTAs per SM) at the cfg64 shape?
Compared with the attention phase of the persistent megakernel and with a plain copy of the same bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ivideogpt_b200 import ops

dev = torch.device("cuda:0")
B, H, Lmax, layers = 64, 12, 752, 12
res = {}
for pos in (514, 632, 750):
    kc = torch.randn(layers, B, H, Lmax, 64, device=dev).to(torch.bfloat16)
    vc = torch.randn(layers, B, H, 64, Lmax, device=dev).to(torch.bfloat16)
    qkv = torch.randn(B, 3 * H * 64, device=dev).to(torch.bfloat16)
    out = torch.empty(B, H * 64, device=dev, dtype=torch.bfloat16)
    cos = torch.randn(1024, 32, device=dev); sin = torch.randn(1024, 32, device=dev)
    for l in range(layers):
        ops.decode_attn_fused(qkv, kc[l], vc[l], out, B, H, Lmax, pos, None, cos, sin, 0.125)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for rep in range(3):
        for l in range(layers):
            ops.decode_attn_fused(qkv, kc[l], vc[l], out, B, H, Lmax, pos, None, cos, sin, 0.125)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (3 * layers)
    byt = B * H * (pos + 1) * 64 * 2 * 2
    res[f"fused_attn_pos{pos}"] = {"us_per_layer": us, "GBps": byt / us / 1e3}
    del kc, vc
x = torch.empty(layers, 124 * 1024 * 1024 // 2, device=dev, dtype=torch.bfloat16)
y = torch.empty_like(x[0])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for l in range(layers):
    y.copy_(x[l])
b.record(); torch.cuda.synchronize()
res["torch_copy_124MB_us"] = a.elapsed_time(b) * 1e3 / layers
print(json.dumps(res))
