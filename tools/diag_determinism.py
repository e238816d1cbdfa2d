"""Which stage is not bit-reproducible / batch-independent?  (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from ivideogpt_b200 import ops
from test_tokenizer import _pair
from oracle.vq_model_ref import TINY_CFG
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)

def same(a, b):
    return bool(torch.equal(a, b)), float((a.float() - b.float()).abs().max())

for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(3, 32, 32, 256, generator=g).to(dt).to(dev)
    s1 = ops.groupnorm_stats(x, 3, 32, 1e-6); s2 = ops.groupnorm_stats(x, 3, 32, 1e-6)
    s1b = ops.groupnorm_stats(x[1:2].contiguous(), 1, 32, 1e-6)
    print(dt, "gn_stats repeat", same(s1, s2), "batch-indep", same(s1[1:2], s1b))
    gam, bet = torch.randn(256, device=dev), torch.randn(256, device=dev)
    y1 = ops.groupnorm_apply(x, s1, gam, bet, True); y2 = ops.groupnorm_apply(x, s1, gam, bet, True)
    print(dt, "gn_apply repeat", same(y1, y2))
    w = (torch.randn(256, 9 * 256, generator=g) * 0.05).to(dt).to(dev)
    b = torch.randn(256, device=dev)
    c1 = ops.conv3x3(x, w, b); c2 = ops.conv3x3(x, w, b); c1b = ops.conv3x3(x[1:2].contiguous(), w, b)
    print(dt, "conv repeat", same(c1, c2), "batch-indep", same(c1[1:2], c1b))
    for bn in (32, 64, 128, 256):
        cb = ops.conv3x3(x, w, b, bn=bn)
        print(dt, "conv bn", bn, same(c1, cb))
_, mine = _pair(TINY_CFG, dev, torch.float32)
tok = torch.cat([torch.randint(0, 512, (3, 513), generator=g), torch.randint(512, 1024, (3, 68), generator=g)], 1)
tok[:, 256] = 1024; tok[:, 513::17] = 1025
tok = tok.to(dev)
f1 = mine.detokenize(tok, 2); f2 = mine.detokenize(tok, 2); o = mine.detokenize(tok[1:2].contiguous(), 2)
print("detok repeat", same(f1, f2), "batch-indep", same(f1[1:2], o))
px = torch.rand(2, 6, 3, 64, 64, generator=g).to(dev)
a = mine.encode_latents(px)[0]; b2 = mine.encode_latents(px)[0]; c = mine.encode_latents(px[:, :2].contiguous(), with_dynamics=False)[0]
print("encode repeat", same(a, b2), "ctx-only", same(a, c))
