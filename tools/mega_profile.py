"""Phase breakdown of the persistent decode megakernel (cycle counters of CTA 0) + wall time vs the CUDA-graph path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_b200_models

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "64"))
tok, llm, _, _ = build_b200_models("cfg64", dev, torch.bfloat16)
eng = llm.b200_engine()
ids = torch.randint(0, 16384, (B, 514), device=dev)
new = 237
res = {}
VARIANTS = (("graph", dict(use_mega=False), 0), ("mega_regs", dict(use_mega=True), 1),
            ("mega_noprefetch", dict(use_mega=True), 2), ("mega", dict(use_mega=True), 0))
if os.environ.get("MEGA_ONLY", "0") == "1":
    VARIANTS = (("mega", dict(use_mega=True), 0),)
for name, kw, mode in VARIANTS:
    eng.mega_attn_mode = mode
    for _ in range(2):
        eng.generate(ids, None, new, True, 100, 1.0, 1, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.generate(ids, None, new, True, 100, 1.0, 1, **kw); b.record(); torch.cuda.synchronize()
    res[name + "_generate_ms"] = a.elapsed_time(b)
# prefill only
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.generate(ids, None, 1, True, 100, 1.0, 1); b.record(); torch.cuda.synchronize()
res["prefill_plus_first_token_ms"] = a.elapsed_time(b)
eng.mega_profile = True
names = ["norm", "qkv", "attention", "o_proj", "gate_up", "down", "lm_head", "sample", "barriers"]
mhz = 1965.0
for mode, tag in ((2, "_noprefetch"), (0, "")):
    eng.mega_attn_mode = mode
    eng.generate(ids, None, new, True, 100, 1.0, 1, use_mega=True)
    torch.cuda.synchronize()
    allc = eng.mega_prof.cpu().tolist()
    cyc = allc[:9]
    res["mega_phase_us_per_step" + tag] = {n: c / (mhz) / (new - 1) for n, c in zip(names, cyc)}
    if mode == 0:
        res["attention_warp0_us_per_step"] = {n: c / mhz / (new - 1) for n, c in
                                              zip(["prologue", "k_loop", "softmax", "v_loop", "tail"], allc[9:14])}
        res["gemm_item_us_per_step"] = {n: c / mhz / (new - 1) for n, c in
                                        zip(["load_a", "slab_wait", "mma_issue", "commit_to_epilogue_end"], allc[14:18])}
res["decode_steps"] = new - 1
print(json.dumps(res))
