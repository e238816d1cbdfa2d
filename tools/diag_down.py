import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from helpers import rel_err
from oracle.llama_ref import TINY_LLAMA, build_hf_llama
from ivideogpt_b200.transformer import B200LlamaForCausalLM
cuda = torch.device("cuda:0")
cfg = dict(TINY_LLAMA, hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_key_value_heads=12)
ref = build_hf_llama(cfg, seed=4321, init_scale=3.0)
mine = B200LlamaForCausalLM(ref.config).to(torch.float32); mine.load_state_dict(ref.state_dict(), strict=True)
mine = mine.to(cuda).eval().set_compute_dtype(torch.bfloat16)
B, L, new, V = 16, 30, 12, 1026
ids = torch.randint(0, V, (B, L), generator=torch.Generator().manual_seed(17)).to(cuda)
eng = mine.b200_engine(); eng.mega_gemm_mode = 0
a = eng.generate(ids, None, new, False, 0, 1.0, 0, use_mega=False)
la = eng.buf("logits", (B, (V + 3) // 4 * 4), torch.float32)[:, :V].clone()
for down in (None, (32, 6), (16, 6)):
    eng.mega_down = down
    m = eng.generate(ids, None, new, False, 0, 1.0, 0, use_mega=True)
    lm = eng.buf("logits", (B, (V + 3) // 4 * 4), torch.float32)[:, :V].clone()
    same = (m == a).all(dim=1)
    print(down, "same rows", int(same.sum()), "rel_err", rel_err(lm[same], la[same]) if int(same.sum()) else None,
          "per-row", [round(rel_err(lm[i], la[i]), 4) for i in range(B) if same[i]][:8])
