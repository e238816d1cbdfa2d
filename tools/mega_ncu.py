"""Short megakernel rollout for an ncu capture (B=64, 514-token prompt, NEW new tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_b200_models
dev = torch.device("cuda:0")
tok, llm, _, _ = build_b200_models("cfg64", dev, torch.bfloat16)
eng = llm.b200_engine()
ids = torch.randint(0, 16384, (64, 514), device=dev)
new = int(os.environ.get("NEW", "33"))
out = eng.generate(ids, None, new, True, 100, 1.0, 1, use_mega=True)
torch.cuda.synchronize()
print("done", out.shape)
