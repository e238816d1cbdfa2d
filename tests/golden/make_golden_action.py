"""Regenerates tests/golden/action_model_tiny.npz by running the REFERENCE's own
ivideogpt/transformer/action_model.py (imported by file path from /root/reference -- it only needs torch and
transformers) around the unmodified HF LlamaForCausalLM with seeded weights.  Run from the repo root in the
build container (the GPU box has no /root/reference; the tests only read the committed .npz):
    python tests/golden/make_golden_action.py
transformers version skew: the reference pins 4.38.2, this image has 5.5.0 (recorded in the fixture)."""
import importlib.util
import os
import sys

import numpy as np
import torch
import transformers

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.action_model_ref import seeded_heads_  # noqa: E402  (only the seeded initialiser is shared)
from oracle.llama_ref import TINY_LLAMA, build_hf_llama  # noqa: E402

REF = "/root/reference/ivideogpt/transformer/action_model.py"
spec = importlib.util.spec_from_file_location("ref_action_model", REF)
ref_mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_mod)

torch.set_num_threads(1)
LAYOUT = dict(action_dim=3, prelude_tokens_num=21, tokens_num_per_dyna=4, context=2, segment_length=5)
LLM_SEED, LLM_SCALE, HEAD_SEED = 991, 4.0, 77
frames = LAYOUT["segment_length"] - LAYOUT["context"]
max_new = frames * (LAYOUT["tokens_num_per_dyna"] + 1) - 1

llm = build_hf_llama(TINY_LLAMA, seed=LLM_SEED, init_scale=LLM_SCALE)
model = ref_mod.HeadModelWithAction(llm, model_type="llama", reward_prediction=True, action_recon=None, **LAYOUT).eval()
seeded_heads_(model, HEAD_SEED)

g = torch.Generator().manual_seed(5)
B = 3
sdf = TINY_LLAMA["vocab_size"] - 1
prompt = torch.randint(0, 1024, (B, LAYOUT["prelude_tokens_num"] + 1), generator=g)
prompt[:, -1] = sdf
action = torch.randn(B, LAYOUT["segment_length"], LAYOUT["action_dim"], generator=g)

# forward (train_gpt.py:792 through action_model.py:154-205): full sequence, labels masked on the prelude
full = torch.randint(0, 1024, (B, LAYOUT["prelude_tokens_num"] + frames * (LAYOUT["tokens_num_per_dyna"] + 1)), generator=g)
for i in range(frames):
    full[:, LAYOUT["prelude_tokens_num"] + i * (LAYOUT["tokens_num_per_dyna"] + 1)] = sdf
labels = full.clone()
labels[:, : LAYOUT["prelude_tokens_num"] + 1] = -100
with torch.no_grad():
    out, reward = model(input_ids=full, labels=labels, action=action)

# generate (greedy): reward_prediction is switched off for the rollout, as mbrl/video_predictor.py does (the reference
# marks its in-generate reward branch as buggy, action_model.py:84-85)
model.reward_prediction = False
gen = model.generate(prompt.clone(), do_sample=False, max_new_tokens=max_new, action=action)
gen_na = model.generate_without_action(prompt.clone(), do_sample=False, max_new_tokens=max_new)

np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "action_model_tiny.npz"),
    layout=np.array([LAYOUT[k] for k in ("action_dim", "prelude_tokens_num", "tokens_num_per_dyna", "context", "segment_length")]),
    seeds=np.array([LLM_SEED, int(LLM_SCALE), HEAD_SEED]), transformers_version=np.array(transformers.__version__),
    prompt=prompt.numpy(), action=action.numpy(), full=full.numpy(), labels=labels.numpy(),
    loss=np.array(float(out.loss)), logits=out.logits.numpy().astype(np.float32), reward=reward.numpy().astype(np.float32),
    generate=gen.numpy(), generate_without_action=gen_na.numpy())
print("loss", float(out.loss), "generate", tuple(gen.shape), "tokens", gen[0, -14:].tolist())
