"""Scratch: gradient deviation of torch's own TF32 GPU arithmetic (cudnn TF32 convs, optionally TF32 matmuls) from the CPU fp32
oracle on the tiny tokenizer training graph -- the yardstick for the sm_100a backward's deviation."""
import os, sys, copy
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle.vq_model_ref import RefCompressiveVQModel, seeded_init_, TINY_CFG

z = np.load("tests/golden/tokenizer_refglue.npz")
px = torch.from_numpy(z["tiny_pixels"])
fut = px.shape[1] - 2
sample, dyn = px[0, :2].contiguous(), px[0, 2:].contiguous()

def grads(model, dev, idx=None):
    model = model.to(dev).train()
    model.zero_grad()
    s, d = sample.to(dev), dyn.to(dev)
    kw = {} if idx is None else dict(idx_ctx=idx[0].to(dev), idx_dyn=idx[1].to(dev))
    out = model.forward_train(s, d, fut, **kw)
    loss = F.mse_loss(out[0], d) + F.mse_loss(out[1], s) + out[2] + 0.5 * out[3]
    loss.backward()
    return {n: p.grad.detach().double().cpu() for n, p in model.named_parameters() if p.grad is not None}

ref = seeded_init_(RefCompressiveVQModel(**TINY_CFG), codebook="normal")
with torch.no_grad():
    zc, zd = ref.encode_latents(px)
    idx = (torch.argmin(torch.cdist(zc, ref.quantize.embedding.weight), 1), torch.argmin(torch.cdist(zd, ref.dynamics_quantize.embedding.weight), 1))
g0 = grads(copy.deepcopy(ref), "cpu", idx)
scale = max(float(v.norm()) for v in g0.values())
for name, conv_tf32, mm_tf32 in (("fp32", False, False), ("tf32conv", True, False), ("tf32all", True, True)):
    torch.backends.cudnn.allow_tf32 = conv_tf32
    torch.backends.cuda.matmul.allow_tf32 = mm_tf32
    g = grads(copy.deepcopy(ref), "cuda", idx)
    errs = sorted(((float((g[n] - g0[n]).norm() / (g0[n].norm() + 1e-4 * scale)), n) for n in g0), reverse=True)
    print(f"== torch cuda {name}: worst {errs[0][0]:.4f} median {errs[len(errs)//2][0]:.5f}")
    for e, n in errs[:25]:
        print(f"   {e:9.5f} {n}")
    np.save(f"gpurun_out/yard_{name}.npy", np.array([(n, e) for e, n in errs], dtype=object), allow_pickle=True)
