"""Action-conditioned wrapper (reference ivideogpt/transformer/action_model.py).

tests/golden/action_model_tiny.npz holds outputs of the REFERENCE FILE ITSELF (imported from /root/reference by
tests/golden/make_golden_action.py, around the unmodified HF Llama, seeded weights).  CPU: the oracle restatement must
reproduce them.  GPU: the B200 product must match them (TF32 path: logits 2e-3, loss 1e-3; greedy tokens equal until a
near-tie of the reference's own logits)."""
import os

import numpy as np
import pytest
import torch

from helpers import ROOT, rel_err

GOLD = os.path.join(ROOT, "tests", "golden", "action_model_tiny.npz")
pytestmark = pytest.mark.usefixtures("deterministic")      # greedy token sequences of two runs are compared exactly


def _gold():
    z = np.load(GOLD)
    keys = ("action_dim", "prelude_tokens_num", "tokens_num_per_dyna", "context", "segment_length")
    layout = dict(zip(keys, (int(v) for v in z["layout"])))
    llm_seed, llm_scale, head_seed = (int(v) for v in z["seeds"])
    return z, layout, llm_seed, float(llm_scale), head_seed


def _oracle(reward=True):
    from oracle.action_model_ref import RefActionModel, seeded_heads_
    from oracle.llama_ref import TINY_LLAMA, build_hf_llama
    z, layout, llm_seed, llm_scale, head_seed = _gold()
    llm = build_hf_llama(TINY_LLAMA, seed=llm_seed, init_scale=llm_scale)
    m = seeded_heads_(RefActionModel(llm, reward_prediction=reward, **layout).eval(), head_seed)
    return z, layout, m


def test_oracle_reproduces_reference_vectors():
    torch.set_num_threads(1)
    z, layout, m = _oracle()
    full, labels, action = (torch.from_numpy(z[k]) for k in ("full", "labels", "action"))
    with torch.no_grad():
        loss, logits, reward = m(full, labels, action)
    assert abs(float(loss) - float(z["loss"])) < 1e-5
    assert rel_err(logits, torch.from_numpy(z["logits"])) < 1e-5
    assert rel_err(reward, torch.from_numpy(z["reward"])) < 1e-5
    frames = layout["segment_length"] - layout["context"]
    max_new = frames * (layout["tokens_num_per_dyna"] + 1) - 1
    prompt = torch.from_numpy(z["prompt"])
    assert np.array_equal(m.generate(prompt.clone(), action, max_new).numpy(), z["generate"])
    assert np.array_equal(m.generate_without_action(prompt.clone(), max_new).numpy(), z["generate_without_action"])


def _product(cuda, dtype, reward=True):
    from ivideogpt_b200.transformer import B200LlamaForCausalLM, HeadModelWithAction
    z, layout, ref = _oracle(reward)
    llm = B200LlamaForCausalLM(ref.llm.config).to(torch.float32)
    llm.load_state_dict(ref.llm.state_dict(), strict=True)
    mine = HeadModelWithAction(llm, model_type="llama", reward_prediction=reward, **layout)
    sd = {k: v for k, v in ref.state_dict().items() if not k.startswith("llm.")}
    missing = mine.load_state_dict(sd, strict=False)
    assert all(k.startswith("llm.") for k in missing.missing_keys) and not missing.unexpected_keys
    mine = mine.to(cuda).eval()
    mine.llm.set_compute_dtype(dtype)
    return z, layout, ref, mine


def _assert_tokens_until_near_tie(got, want, ref, action, layout, tol=2e-2):
    """Rows may leave the reference's greedy path only where the reference's own top-2 logit margin is tiny."""
    T0 = layout["prelude_tokens_num"] + 1
    for b in range(want.shape[0]):
        neq = (got[b] != want[b]).nonzero()
        if len(neq) == 0:
            continue
        p = int(neq[0])
        assert p >= T0, "prompt tokens were modified"
        seq = torch.cat([want[b:b + 1, :p], torch.zeros(1, want.shape[1] + 1 - p, dtype=torch.int64)], dim=1)
        with torch.no_grad():
            lg = ref(seq, None, action[b:b + 1])[1][0, p - 1] if action is not None else \
                ref.llm(input_ids=want[b:b + 1, :p]).logits[0, -1]
        top2 = lg.topk(2).values
        assert float(top2[0] - top2[1]) < tol * float(lg.abs().max()), f"row {b} diverged at {p} without a near-tie"


@pytest.mark.gpu
def test_forward_vs_reference_vectors(cuda):
    z, layout, ref, mine = _product(cuda, torch.float32)
    full, labels, action = (torch.from_numpy(z[k]).to(cuda) for k in ("full", "labels", "action"))
    with torch.no_grad():
        out, reward = mine(input_ids=full, labels=labels, action=action)
    assert rel_err(out.logits, torch.from_numpy(z["logits"])) < 2e-3
    assert abs(float(out.loss) - float(z["loss"])) / float(z["loss"]) < 1e-3
    assert rel_err(reward, torch.from_numpy(z["reward"])) < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("persistent", [False, True])
def test_generate_vs_reference_vectors(cuda, persistent):
    """persistent=False re-prefills every frame like the reference; True keeps one KV cache for the whole rollout
    (SURVEY 8f rank 1): same tokens, one prefill."""
    z, layout, ref, mine = _product(cuda, torch.float32, reward=False)
    mine.persistent_cache = persistent
    frames = layout["segment_length"] - layout["context"]
    max_new = frames * (layout["tokens_num_per_dyna"] + 1) - 1
    prompt, action = torch.from_numpy(z["prompt"]), torch.from_numpy(z["action"])
    got = mine.generate(prompt.to(cuda), do_sample=False, max_new_tokens=max_new, action=action.to(cuda)).cpu()
    want = torch.from_numpy(z["generate"])
    assert got.shape == want.shape
    _assert_tokens_until_near_tie(got, want, ref, action, layout)
    got = mine.generate_without_action(prompt.to(cuda), do_sample=False, max_new_tokens=max_new).cpu()
    _assert_tokens_until_near_tie(got, torch.from_numpy(z["generate_without_action"]), ref, None, layout)


@pytest.mark.gpu
def test_persistent_rollout_bf16_megakernel_matches_reprefill(cuda):
    """bf16: the one-launch persistent rollout (decode megakernel with forced separator slots and per-slot action
    embeddings) against the per-frame re-prefill path on the same weights; separators sit where the layout puts them."""
    z, layout, ref, mine = _product(cuda, torch.bfloat16, reward=False)
    frames = layout["segment_length"] - layout["context"]
    n = layout["tokens_num_per_dyna"]
    max_new = frames * (n + 1) - 1
    prompt, action = torch.from_numpy(z["prompt"]).to(cuda), torch.from_numpy(z["action"]).to(cuda)
    mine.persistent_cache = False
    a = mine.generate(prompt, do_sample=False, max_new_tokens=max_new, action=action).cpu()
    mine.persistent_cache = True
    b = mine.generate(prompt, do_sample=False, max_new_tokens=max_new, action=action).cpu()
    assert a.shape == b.shape
    sdf = ref.sdf
    slots = [layout["prelude_tokens_num"] + i * (n + 1) for i in range(frames)]
    assert all(bool((b[:, s] == sdf).all()) for s in slots)
    agree = (a == b).float().mean().item()
    assert agree > 0.9, f"persistent vs re-prefill token agreement {agree}"   # bf16 near-ties may flip single tokens
    # sampling mode: shapes, separators and vocabulary range
    s = mine.generate(prompt, do_sample=True, top_k=50, max_new_tokens=max_new, action=action).cpu()
    assert s.shape == a.shape and all(bool((s[:, k] == sdf).all()) for k in slots) and int(s.max()) <= sdf


def test_persistent_rollout_layout_matches_the_reference_append_positions():
    """CPU, host logic only: the forced-separator layout handed to the decode kernels must reproduce where the reference's
    loop puts separators (one after every `per_frame` generated tokens, action_model.py:109-110) and where it adds action
    embeddings (prelude + i * (n + 1), :80-81); layouts the single-cache rollout cannot express fall back (None)."""
    from types import SimpleNamespace
    from ivideogpt_b200.transformer import HeadModelWithAction
    llm = torch.nn.Module()
    llm.config = SimpleNamespace(vocab_size=1026, hidden_size=128)
    llm.b200_engine = lambda: None
    P, n, ctx, seg = 21, 4, 2, 5
    m = HeadModelWithAction(llm, action_dim=3, prelude_tokens_num=P, tokens_num_per_dyna=n, context=ctx, segment_length=seg)
    frames = seg - ctx
    max_new = frames * (n + 1) - 1
    per_frame = (max_new + 1) // frames - 1
    assert per_frame == n
    T = P + 1
    # reference: separators land at T + per_frame + j * (per_frame + 1); action slots at P + i * (n + 1)
    ref_sep = [T + per_frame + j * (per_frame + 1) for j in range(frames - 1)]
    slot0, period = m._persistent_layout(T, per_frame, True)
    assert (slot0, period) == (P, n + 1)
    assert [q for q in range(T, T + max_new) if (q - slot0) % period == 0] == ref_sep
    assert [slot0 + i * period for i in range(frames)] == [P + i * (n + 1) for i in range(frames)]
    s0, per = m._persistent_layout(T, per_frame, False)            # generate_without_action: separators only
    assert [q for q in range(T, T + max_new) if q >= s0 and (q - s0) % per == 0] == ref_sep
    assert m._persistent_layout(T + 2, per_frame, True) is None     # prompt does not end on the first separator slot
    assert m._persistent_layout(T, per_frame + 1, True) is None     # frame length differs from tokens_num_per_dyna
    m.persistent_cache = False
    assert m._persistent_layout(T, per_frame, True) is None


@pytest.mark.skipif(not os.path.isfile("/root/reference/ivideogpt/transformer/action_model.py"),
                    reason="needs the reference checkout (build container only)")
def test_state_dict_keys_equal_the_reference_own_class():
    """Parameter names / shapes of HeadModelWithAction (with reward and action-reconstruction heads) against the reference's
    own class around the HF Llama: released action-conditioned checkpoints load with strict=True."""
    import importlib.util
    from ivideogpt_b200.transformer import B200LlamaForCausalLM, HeadModelWithAction
    from oracle.llama_ref import TINY_LLAMA, build_hf_llama
    spec = importlib.util.spec_from_file_location("ref_action_model_t", "/root/reference/ivideogpt/transformer/action_model.py")
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    hf = build_hf_llama(TINY_LLAMA)
    kw = dict(action_dim=3, prelude_tokens_num=21, tokens_num_per_dyna=4, context=2, segment_length=5,
              model_type="llama", reward_prediction=True, action_recon=0.1)
    ref = ref_mod.HeadModelWithAction(hf, **kw)
    mine = HeadModelWithAction(B200LlamaForCausalLM(hf.config), **kw)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert a == b, sorted(set(a.items()) ^ set(b.items()))[:6]
    assert mine.token_for_sdf == ref.token_for_sdf == 1025
    with pytest.raises(ValueError):
        HeadModelWithAction(B200LlamaForCausalLM(hf.config), **dict(kw, model_type="gpt2"))
