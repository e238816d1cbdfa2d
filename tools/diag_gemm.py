"""GEMM bring-up diagnostics: structured inputs whose wrong outputs reveal WHICH layout assumption is broken."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ivideogpt_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def report(name, got, want):
    got, want = got.double().cpu(), want.double().cpu()
    err = (got - want).norm() / want.norm()
    bad = (got - want).abs() > 1e-3 * want.abs().max()
    print(f"[{name}] rel_err={err:.3e} bad={int(bad.sum())}/{bad.numel()} nan={int(torch.isnan(got).sum())}")
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()[:8].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        print("   bad rows(first 8):", rows, " bad cols(first 16):", cols)
        r, c = (rows[0] if rows else 0), (cols[0] if cols else 0)
        print("   got [r, :8]", got[r, :8].tolist())
        print("   want[r, :8]", want[r, :8].tolist())


for dtype in (torch.bfloat16, torch.float32):
    for (M, N, K, bn) in [(128, 128, 64, 128), (128, 128, 128, 128), (128, 32, 64, 32), (256, 128, 256, 128)]:
        a = torch.randint(-2, 3, (M, K)).to(dtype).to(dev)
        w = torch.randint(-2, 3, (N, K)).to(dtype).to(dev)
        try:
            out = ops.gemm(a, w, out_dtype=torch.float32, bn=bn)
            torch.cuda.synchronize()
            report(f"{dtype} M{M} N{N} K{K} bn{bn} int-valued", out, a.double() @ w.double().t())
        except Exception as e:  # noqa
            print(f"[{dtype} M{M} N{N} K{K}] EXC {e}")
    # identity probes: A = I (K = M = 64 padded) -> out should equal W^T block
    M = N = 128; K = 64 if dtype == torch.bfloat16 else 32
    a = torch.zeros(M, K); a[:K, :K] = torch.eye(K)
    w = torch.arange(N * K, dtype=torch.float32).reshape(N, K) % 251
    out = ops.gemm(a.to(dtype).to(dev), w.to(dtype).to(dev), out_dtype=torch.float32, bn=128)
    torch.cuda.synchronize()
    report(f"{dtype} identity-A", out, a.to(dtype).double() @ w.to(dtype).double().t())
print("diag done")
