import os, sys, torch, traceback
sys.path.insert(0, os.getcwd())
from bench import build_b200_models, synthetic_clips
dev = torch.device("cuda:0")
tok, llm, _, _ = build_b200_models("cfg64", dev, torch.bfloat16)
clips = synthetic_clips(64, 16, 64).to(dev)
stage = "?"
try:
    for it in range(3):
        stage = "tokenize"; prompt = tok.tokenize_context(clips); torch.cuda.synchronize()
        stage = "generate"; out = llm.generate(prompt, do_sample=False, max_new_tokens=20); torch.cuda.synchronize()
        full = torch.cat([out, torch.randint(8192, 16384, (64, 751 - out.shape[1]), device=dev)], 1)
        full[:, 513::17] = 16385
        stage = "detokenize"; tok.detokenize(full, 2); torch.cuda.synchronize()
        stage = "tokenize-all"; tok.tokenize(clips, 2); torch.cuda.synchronize()
    print("all ok")
except Exception as e:
    print("FAILED in", stage, repr(e)[:200])
