"""Representative conv launches of the cfg64 / cfg256 decoders for ncu (`ncu --set full -k regex:gemm_tc_kernel`), and a CUDA-event
timing of the same launches.  Shapes: (frames, H, W, Cin, Cout) of the up path (3x3, stride 1, bf16)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ivideogpt_b200 import ops

dev = torch.device("cuda:0")
shapes = [(896, 64, 64, 128, 128), (896, 32, 32, 256, 256), (896, 16, 16, 512, 512), (896, 64, 64, 256, 128),
          (224, 256, 256, 128, 128), (224, 64, 64, 256, 256)]
if os.environ.get("SHAPES"):
    shapes = [shapes[int(i)] for i in os.environ["SHAPES"].split(",")]
res = []
for (N, H, W, Cin, Cout) in shapes:
    x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device=dev) * 0.02).to(torch.bfloat16)
    b = torch.zeros(Cout, device=dev)
    r = torch.randn(N, H, W, Cout, device=dev).to(torch.bfloat16)
    for _ in range(2):
        y = ops.conv3x3(x, w, b, residual=r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = ops.conv3x3(x, w, b, residual=r)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * N * H * W * Cout * 9 * Cin
    res.append({"shape": [N, H, W, Cin, Cout], "ms": ms, "TFLOPs": fl / ms / 1e9})
    del x, w, r, y
print(json.dumps(res))
