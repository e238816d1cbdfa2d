"""Action-conditioned rollout (HeadModelWithAction.generate, reference action_model.py:56-121): one persistent KV cache
(this repo) against the reference-shaped loop that re-prefills the whole history for every frame.
Same weights, same kernels; prints rollout ms and predicted frames/s for both."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_b200_models
from ivideogpt_b200.transformer import HeadModelWithAction

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "64"))
ctx, seg, n = 2, 16, 16
tok, llm, _, _ = build_b200_models("cfg64", dev, torch.bfloat16)
P = ctx * 257 - 1
model = HeadModelWithAction(llm, action_dim=4, prelude_tokens_num=P, tokens_num_per_dyna=n, context=ctx, segment_length=seg).to(dev).eval()
torch.nn.init.normal_(model.action_linear.weight, std=0.02)
g = torch.Generator().manual_seed(0)
prompt = torch.randint(0, 8192, (B, P + 1), generator=g).to(dev)
prompt[:, -1] = model.token_for_sdf
action = torch.randn(B, seg, 4, generator=g).to(dev)
max_new = (seg - ctx) * (n + 1) - 1
res = {"B": B, "prompt": P + 1, "max_new_tokens": max_new}
for name, persistent in (("persistent_cache", True), ("reprefill_per_frame", False)):
    model.persistent_cache = persistent
    for _ in range(2):
        out = model.generate(prompt, do_sample=True, top_k=100, max_new_tokens=max_new, action=action)
    assert out.shape == (B, P + 1 + max_new)
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model.generate(prompt, do_sample=True, top_k=100, max_new_tokens=max_new, action=action); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[1]
    res[name] = {"rollout_ms": ms, "predicted_frames_per_s": B * (seg - ctx) / (ms / 1e3)}
res["speedup"] = res["reprefill_per_frame"]["rollout_ms"] / res["persistent_cache"]["rollout_ms"]
print(json.dumps(res))
