"""VQ-argmin kernel timing (BASELINE metric "VQ-argmin HBM GB/s"): CUDA events around the launch sequence,
L2 flushed between iterations.  Reports algorithmic GB/s (4ND + 4KD + 8N bytes) and fp32 TFLOP/s (2NKD)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ivideogpt_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
from ivideogpt_b200 import _lib
res = []
for order in (0, 1):
  _lib.load().ivgpt_vq_set_order(order)
  for name, N in (("cfg64_ctx", 32768), ("cfg64_dyn", 14336), ("cfg256_ctx", 8192), ("cfg256_dyn", 3584)):
      K, D = 8192, 64
      z = torch.randn(N, D, device=dev)
      e = torch.randn(K, D, device=dev)
      for _ in range(3):
          ops.vq_argmin(z, e)
      ts = []
      for _ in range(10):
          flush.zero_()
          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a.record(); ops.vq_argmin(z, e); b.record(); torch.cuda.synchronize()
          ts.append(a.elapsed_time(b))
      t = sorted(ts)[len(ts) // 2] * 1e-3
      byt = 4 * N * D + 4 * K * D + 8 * N
      res.append({"order": order, "case": name, "N": N, "ms": t * 1e3, "alg_GBps": byt / t / 1e9, "fp32_TFLOPs": 2.0 * N * K * D / t / 1e12})
print(json.dumps(res))
