"""Phase breakdown of the persistent decode megakernel (cycle counters of CTA 0) + wall time per operand-path variant.
VARIANTS env: comma list of gemm_mode:bn_wide:unused[:attn_mode[:qkv_splits:o_splits:d_splits]] (bn_wide 0/1 = automatic),
default "0:0:1"."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_b200_models

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "64"))
workload = os.environ.get("WORKLOAD", "cfg64")
tok, llm, _, _ = build_b200_models(workload, dev, torch.bfloat16)
eng = llm.b200_engine()
ids = torch.randint(0, 16384, (B, 514), device=dev)
new = 237
mhz = 1965.0
names = ["norm", "qkv", "attention", "o_proj", "gate_up", "down", "lm_head", "sample", "barriers"]
res = {"B": B, "workload": workload, "decode_steps": new - 1}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.generate(ids, None, 1, True, 100, 1.0, 1)
a.record(); eng.generate(ids, None, 1, True, 100, 1.0, 1); b.record(); torch.cuda.synchronize()
res["prefill_plus_first_token_ms"] = a.elapsed_time(b)
if os.environ.get("WITH_GRAPH", "0") == "1":
    for _ in range(2):
        eng.generate(ids, None, new, True, 100, 1.0, 1, use_mega=False)
    a.record(); eng.generate(ids, None, new, True, 100, 1.0, 1, use_mega=False); b.record(); torch.cuda.synchronize()
    res["graph_generate_ms"] = a.elapsed_time(b)
for spec in os.environ.get("VARIANTS", "0:0:1").split(","):
    f = [int(v) for v in spec.split(":")]
    eng.mega_gemm_mode = f[0]
    eng.mega_bn_wide = f[1] if f[1] >= 16 else 0
    eng.mega_down = (f[2] // 100, f[2] % 100) if f[2] >= 100 else None      # third field: 100 * bn_down + d_splits (e.g. 3206)
    eng.mega_attn_mode = f[3] if len(f) > 3 else 0
    eng.mega_splits_override = tuple(f[4:7]) if len(f) >= 7 else None
    tag = f"gemm={f[0]},bn_wide={eng.mega_bn_wide or 'auto'},down={eng.mega_down},attn={eng.mega_attn_mode},splits={eng.mega_splits_override}"
    PROFILE = os.environ.get("PHASES", "1") == "1"
    eng.mega_profile = False
    try:
        for _ in range(2):
            eng.generate(ids, None, new, True, 100, 1.0, 1, use_mega=True)
        ts = []
        for _ in range(3):
            a.record(); eng.generate(ids, None, new, True, 100, 1.0, 1, use_mega=True); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[tag] = {"generate_ms": sorted(ts)[1], "generate_ms_all": ts}
        if PROFILE:
            eng.mega_profile = True
            eng.generate(ids, None, new, True, 100, 1.0, 1, use_mega=True)
            torch.cuda.synchronize()
            allc = eng.mega_prof.cpu().tolist()
            res[tag].update({
                "phase_us_per_step": {n: round(c / mhz / (new - 1), 2) for n, c in zip(names, allc[:9])},
                "attention_warp0_us_per_step": {n: round(c / mhz / (new - 1), 2) for n, c in
                                                zip(["prologue", "k_loop", "softmax", "v_loop", "tail"], allc[9:14])},
                "gemm_item_us_per_step": {n: round(c / mhz / (new - 1), 2) for n, c in
                                          zip(["load_a", "slab_wait", "mma_issue", "commit_to_epilogue_end", "gemm_phase_total"], allc[14:19])}})
    except Exception as e:  # a variant that faults must not hide the others
        res[tag] = {"error": repr(e)[:300]}
        break
print(json.dumps(res))
