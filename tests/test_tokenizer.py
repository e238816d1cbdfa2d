"""Tokenizer half: B200 CompressiveVQModel vs the CPU fp32 oracle restatement on shared seeded weights."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import ROOT, rel_err


def _pair(cfg, cuda=None, dtype=torch.float32, codebook="normal"):
    from oracle.vq_model_ref import RefCompressiveVQModel, seeded_init_
    from ivideogpt_b200.vq_model import CompressiveVQModel
    ref = seeded_init_(RefCompressiveVQModel(**cfg).eval(), codebook=codebook)
    mine = CompressiveVQModel.from_config(cfg)
    mine.load_state_dict(ref.state_dict(), strict=True)
    if cuda is not None:
        mine = mine.to(cuda).eval().set_compute_dtype(dtype)
    return ref, mine


def _cfg(name):
    with open(os.path.join(ROOT, "configs", name + ".json")) as fh:
        return {k: v for k, v in json.load(fh).items() if not k.startswith("_")}


def test_state_dict_layout_and_param_counts():
    """Key names / shapes equal the oracle's (which mirror the reference module tree) and the README counts."""
    for name, want in (("ctx_vae64", 114.2), ("ctx_vae256", 310.5)):
        ref, mine = _pair(_cfg(name))
        a, b = mine.state_dict(), ref.state_dict()
        assert set(a) == set(b)
        assert all(a[k].shape == b[k].shape for k in a)
        assert abs(sum(p.numel() for p in mine.parameters()) / 1e6 - want) < 0.06
    assert mine.context_length == 2 and mine.num_vq_embeddings == 8192 and mine.num_dyn_embeddings == 8192
    assert any("quantize" in n for n, _ in mine.named_parameters())        # train_tokenizer.py:424 name filter
    assert mine.cond_decoder.conv_out.weight.shape == (3, 128, 3, 3)        # train_tokenizer.py:714 attribute path


def test_save_load_roundtrip_and_cpu_refusal(tmp_path):
    from ivideogpt_b200.vq_model import CompressiveVQModel
    from oracle.vq_model_ref import TINY_CFG
    _, mine = _pair(TINY_CFG)
    mine.save_pretrained(str(tmp_path / "tokenizer"))
    again = CompressiveVQModel.from_pretrained(str(tmp_path), subfolder="tokenizer", low_cpu_mem_usage=False)
    assert all(torch.equal(v, again.state_dict()[k]) for k, v in mine.state_dict().items())
    assert again.config["context_length"] == 2 and again.config.patch_size == 4
    with pytest.raises(RuntimeError):
        again.tokenize(torch.rand(1, 4, 3, 64, 64), 2)                      # CPU: loud failure, no fallback
    again.set_context_length(1)
    assert again.cond_encoder.cross_att_blocks[0].kv_pos_emb.shape[0] == 256


def test_oracle_golden_vectors():
    """The committed golden fixture (made by tests/golden/make_golden.py from the oracle) still reproduces."""
    from oracle.vq_model_ref import TINY_CFG
    path = os.path.join(ROOT, "tests", "golden", "tokenizer_tiny.npz")
    gold = np.load(path)
    ref, _ = _pair(TINY_CFG)
    px = torch.from_numpy(gold["pixels"])
    tok, lab = ref.tokenize(px, 2)
    # BLAS summation order differs between hosts: allow isolated near-tie flips, nothing structural
    assert (tok.numpy() == gold["tokens"]).mean() > 0.99
    assert np.array_equal(lab.numpy() == -100, gold["labels"] == -100)
    rec = ref.detokenize(torch.from_numpy(gold["tokens"]), 2)
    assert rel_err(rec, torch.from_numpy(gold["recon"])) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol_lat,tol_px,min_match", [(torch.float32, 2e-3, 3e-3, 0.97),
                                                            (torch.bfloat16, 3e-2, 4e-2, 0.80)])
def test_tiny_tokenizer_vs_oracle(cuda, dtype, tol_lat, tol_px, min_match):
    from oracle.vq_model_ref import TINY_CFG
    ref, mine = _pair(TINY_CFG, cuda, dtype)
    px = torch.rand(2, 6, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    zc_ref, zd_ref = ref.encode_latents(px)
    zc, zd = mine.encode_latents(px.to(cuda))
    assert rel_err(zc, zc_ref) < tol_lat and rel_err(zd, zd_ref) < tol_lat
    tok_ref, lab_ref = ref.tokenize(px, 2)
    tok, lab = mine.tokenize(px.to(cuda), 2)
    assert tok.shape == tok_ref.shape and tok.dtype == torch.int64
    match = (tok.cpu() == tok_ref).float().mean().item()
    assert match >= min_match, f"token agreement {match:.4f}"
    # structural positions (separators, -100 labels) are exact regardless of float noise
    sep = tok_ref >= 1024
    assert torch.equal(tok.cpu()[sep], tok_ref[sep]) and torch.equal(lab.cpu() == -100, lab_ref == -100)
    # detokenize from the ORACLE's tokens so both sides decode the same ids
    rec_ref = ref.detokenize(tok_ref, 2)
    rec = mine.detokenize(tok_ref.to(cuda), 2)
    assert rec.shape == rec_ref.shape and rec.dtype == torch.float32
    assert rel_err(rec, rec_ref) < tol_px
    # golden fixture
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tokenizer_tiny.npz"))
    rec_g = mine.detokenize(torch.from_numpy(gold["tokens"]).to(cuda), 2)
    assert rel_err(rec_g, torch.from_numpy(gold["recon"])) < tol_px


@pytest.mark.gpu
@pytest.mark.usefixtures("deterministic")      # tokenize() and tokenize_context() of the same clip are compared exactly
def test_cfg64_tokenizer_vs_oracle(cuda):
    """BASELINE config ctx_vae64 (114 M), one 64x64x16 clip, TF32 path."""
    ref, mine = _pair(_cfg("ctx_vae64"), cuda, torch.float32)
    px = torch.rand(1, 16, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    zc_ref, zd_ref = ref.encode_latents(px)
    zc, zd = mine.encode_latents(px.to(cuda))
    assert rel_err(zc, zc_ref) < 3e-3 and rel_err(zd, zd_ref) < 3e-3
    tok_ref, _ = ref.tokenize(px, 2)
    tok, _ = mine.tokenize(px.to(cuda), 2)
    assert tok.shape == (1, 751)
    assert (tok.cpu() == tok_ref).float().mean().item() > 0.95
    rec_ref = ref.detokenize(tok_ref, 2)
    rec = mine.detokenize(tok_ref.to(cuda), 2)
    assert rec.shape == (1, 16, 3, 64, 64)
    assert rel_err(rec, rec_ref) < 5e-3
    ctx_only = mine.tokenize_context(px.to(cuda))
    assert torch.equal(ctx_only, tok[:, :514])


@pytest.mark.gpu
@pytest.mark.usefixtures("deterministic")
def test_detokenize_batch_independence_and_cache(cuda):
    """Size-independent properties: clips are independent units (batching must not change any clip), and the
    cached-context path reproduces the uncached frames."""
    from oracle.vq_model_ref import TINY_CFG
    _, mine = _pair(TINY_CFG, cuda, torch.float32)
    g = torch.Generator().manual_seed(4)
    tok = torch.cat([torch.randint(0, 512, (3, 513), generator=g), torch.randint(512, 1024, (3, 68), generator=g)], 1)
    tok[:, 256] = 1024
    tok[:, 513::17] = 1025
    tok = tok.to(cuda)
    full = mine.detokenize(tok, 2)
    one = mine.detokenize(tok[1:2].contiguous(), 2)
    assert rel_err(one, full[1:2]) < 1e-6
    rec, cache = mine.detokenize(tok, 2, return_cache=True)
    again = mine.detokenize(tok, 2, cache=cache)
    assert torch.equal(rec, full) and rel_err(again, full) < 1e-6


def test_oracle_reproduces_vectors_of_the_reference_own_tokenizer_code():
    """tests/golden/tokenizer_refglue.npz was produced by RUNNING the reference's own vae.py / conditional_vae.py /
    compressive_vq_model.py (unmodified, imported from /root/reference) on top of oracle/diffusers_stub (stand-ins for the
    diffusers 0.27.0 building blocks, which cannot be installed offline) -- tests/golden/make_golden_tokenizer_ref.py.
    The oracle must reproduce those vectors: tokens and labels exactly, pixels to fp32 rounding.  (At generation time the
    same comparison also ran for the full ctx_vae256 model; its summary is stored in the fixture.)"""
    import json
    import numpy as np
    from oracle.vq_model_ref import TINY_CFG, RefCompressiveVQModel, config_path, seeded_init_
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    z = np.load(os.path.join(ROOT, "tests", "golden", "tokenizer_refglue.npz"))
    summary = json.loads(str(z["summary"]))
    assert set(summary) == {"tiny", "cfg64", "cfg256_not_stored"}
    assert all(d["tokens_equal"] and d["labels_equal"] and d["recon_max_abs_diff"] < 1e-5 for d in summary.values())
    with open(config_path("ctx_vae64")) as fh:
        cfg64 = {k: v for k, v in json.load(fh).items() if not k.startswith("_") and k not in ("down_block_types", "up_block_types")}
    for name, cfg in (("tiny", TINY_CFG), ("cfg64", cfg64)):
        oracle = seeded_init_(RefCompressiveVQModel(**cfg).eval())
        px = torch.from_numpy(z[f"{name}_pixels"])
        with torch.no_grad():
            tok, lab = oracle.tokenize(px, cfg["context_length"])
            rec = oracle.detokenize(tok, cfg["context_length"])
        assert np.array_equal(tok.numpy(), z[f"{name}_tokens"]) and np.array_equal(lab.numpy(), z[f"{name}_labels"]), name
        assert float(np.abs(rec.numpy() - z[f"{name}_recon"]).max()) < 1e-4, name
    assert z["tiny_recon_cached"].shape == (1, 3, 3, 64, 64)      # context (2) + the one future frame decoded from the cache
    # the training graph (forward() :332-369 + decode() :290-330): values and gradients of the reference's own forward
    assert summary["tiny"]["train_forward_max_abs_diff"] < 1e-5 and summary["tiny"]["train_grad_max_rel_diff"] < 1e-4
    oracle = seeded_init_(RefCompressiveVQModel(**TINY_CFG).eval())
    px = torch.from_numpy(z["tiny_pixels"])
    outs = oracle.forward_train(px[0, :2].clone(), px[0, 2:].clone(), px.shape[1] - 2)
    assert float(np.abs(outs[0].detach().numpy() - z["tiny_train_dec"]).max()) < 1e-4
    assert float(np.abs(outs[1].detach().numpy() - z["tiny_train_ref_dec"]).max()) < 1e-4
    assert np.allclose([float(outs[2]), float(outs[3])], z["tiny_train_losses"], rtol=1e-4)
    (sum(((x - 0.5) ** 2).mean() for x in outs[:2]) + outs[2] + outs[3]).backward()
    norms = [float(p.grad.norm()) for p in (oracle.quant_conv.weight, oracle.cond_decoder.conv_out.weight,
                                            oracle.quantize.embedding.weight, oracle.encoder.conv_in.weight)]
    assert np.allclose(norms, z["tiny_train_grad_norms"], rtol=1e-3)


def test_legacy_checkpoint_keys_bin_fallback_and_mismatched_sizes(tmp_path):
    """Host-side checkpoint I/O (hub_io.py): old diffusers attention key names are converted on load, the .bin fallback is
    read when no safetensors file exists, ignore_mismatched_sizes drops tensors whose shape changed."""
    from ivideogpt_b200.vq_model import CompressiveVQModel
    from ivideogpt_b200.vq_model.hub_io import WEIGHTS_BIN
    from oracle.vq_model_ref import TINY_CFG
    _, mine = _pair(TINY_CFG)
    sd = mine.state_dict()
    ren = {".to_q.": ".query.", ".to_k.": ".key.", ".to_v.": ".value.", ".to_out.0.": ".proj_attn."}
    legacy = {}
    for k, v in sd.items():
        if ".attentions." in k:
            for a, b in ren.items():
                k = k.replace(a, b)
        legacy[k] = v.clone()
    assert any(".query." in k for k in legacy), "the tiny config has a mid-block attention (cond encoder / decoder)"
    d = tmp_path / "tokenizer"
    mine.save_pretrained(str(d), save_function=lambda state, path: torch.save(legacy, path))     # writes the .bin name
    assert os.path.exists(d / WEIGHTS_BIN) and not os.path.exists(d / "diffusion_pytorch_model.safetensors")
    again = CompressiveVQModel.from_pretrained(str(tmp_path), subfolder="tokenizer")
    assert all(torch.equal(v, again.state_dict()[k]) for k, v in sd.items())
    # a codebook of another size: strict load refuses, ignore_mismatched_sizes keeps the model's own tensor
    legacy["quantize.embedding.weight"] = torch.zeros(7, legacy["quantize.embedding.weight"].shape[1])
    torch.save(legacy, d / WEIGHTS_BIN)
    with pytest.raises(RuntimeError):
        CompressiveVQModel.from_pretrained(str(tmp_path), subfolder="tokenizer")
    loose = CompressiveVQModel.from_pretrained(str(tmp_path), subfolder="tokenizer", ignore_mismatched_sizes=True)
    assert loose.state_dict()["quantize.embedding.weight"].shape == sd["quantize.embedding.weight"].shape
    assert torch.equal(loose.state_dict()["quant_conv.weight"], sd["quant_conv.weight"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/ivideogpt/vq_model"), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("config", ["ctx_vae64", "ctx_vae256"])
def test_module_tree_equals_the_reference_own_class(config):
    """The product's parameter names and shapes against the REFERENCE'S OWN CompressiveVQModel class (its three files run on
    oracle/diffusers_stub) for the released configs: a released checkpoint loads with strict=True."""
    import importlib.util
    import sys
    from ivideogpt_b200.vq_model import CompressiveVQModel
    from oracle.vq_model_ref import config_path
    stub = os.path.join(ROOT, "oracle", "diffusers_stub")
    sys.path.insert(0, stub)
    try:
        ref_dir = "/root/reference/ivideogpt/vq_model"
        spec = importlib.util.spec_from_file_location("ref_vq_model_t", ref_dir + "/__init__.py", submodule_search_locations=[ref_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["ref_vq_model_t"] = pkg
        spec.loader.exec_module(pkg)
        with open(config_path(config)) as fh:
            cfg = {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
        with torch.device("meta"):
            ref = pkg.CompressiveVQModel(**cfg)
            mine = CompressiveVQModel.from_config(cfg)
        a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
        assert a == b, sorted(set(a.items()) ^ set(b.items()))[:6]
        # README.md:53-57 of the reference: 114.16 M / 310.47 M tokenizer parameters
        assert sum(int(np.prod(s)) for s in a.values()) == {"ctx_vae64": 114_159_174, "ctx_vae256": 310_472_774}[config]
    finally:
        sys.path.remove(stub)
        for k in [k for k in sys.modules if k == "diffusers" or k.startswith("diffusers.") or k.startswith("ref_vq_model_t")]:
            del sys.modules[k]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol,tol_loss", [(torch.float32, 4e-3, 1e-2), (torch.bfloat16, 4e-2, 6e-2)])
@pytest.mark.usefixtures("deterministic")      # forward() twice and encode_latents(): identical VQ choices are assumed
def test_training_graph_forward_vs_oracle(cuda, dtype, tol, tol_loss):
    """Row f3 (forward half): CompressiveVQModel.forward(sample=, dyn_sample=, segment_len=) against the oracle's forward_train
    -- itself pinned, values AND gradients, to the reference's own forward() (tests/golden/tokenizer_refglue.npz).  The graph
    contains two discrete choices (VQ argmin); rounding may flip one on a near-tie and a flipped code changes a whole latent
    patch, so the comparison holds the choice fixed: the product's own indices (checked to differ from the oracle's only on
    near-ties) are handed to the oracle, then both reconstructions and both commit losses must agree.
    The backward half: test_training_graph_backward_vs_oracle."""
    from ivideogpt_b200 import ops
    from oracle.vq_model_ref import TINY_CFG
    z = np.load(os.path.join(ROOT, "tests", "golden", "tokenizer_refglue.npz"))
    ref, mine = _pair(TINY_CFG, cuda, dtype)
    px = torch.from_numpy(z["tiny_pixels"])
    fut = px.shape[1] - 2
    sample, dyn = px[0, :2].contiguous(), px[0, 2:].contiguous()
    with torch.no_grad():
        dec, ref_dec, commit, dyn_commit = mine(sample=sample.to(cuda), dyn_sample=dyn.to(cuda), segment_len=fut,
                                                return_dict=False, return_loss=True)
        rec = mine(sample=sample.to(cuda), dyn_sample=dyn.to(cuda), segment_len=fut, return_loss=True)
        zc, zd = mine.encode_latents(px.to(cuda))                      # the same arithmetic as inside forward()
        idx_c = ops.vq_argmin(zc, mine.quantize.embedding.weight.detach().float()).cpu()
        idx_d = ops.vq_argmin(zd, mine.dynamics_quantize.embedding.weight.detach().float()).cpu()
        zc_ref, zd_ref = ref.encode_latents(px)
        flips = 0
        for z_ref, z_mine, idx, cb in ((zc_ref, zc, idx_c, ref.quantize.embedding.weight),
                                       (zd_ref, zd, idx_d, ref.dynamics_quantize.embedding.weight)):
            d = torch.cdist(z_ref.double(), cb.detach().double())
            best = d.min(1).values
            ours = d.gather(1, idx[:, None]).squeeze(1)
            dz = (z_mine.cpu().double() - z_ref.double()).norm(dim=1)
            assert bool((ours - best <= 2.0 * dz + 1e-9).all()), "a VQ index differs from the oracle's without a near-tie"
            flips += int((idx != d.argmin(1)).sum())
        want = ref.forward_train(sample, dyn, fut, idx_ctx=idx_c, idx_dyn=idx_d)
    print(f"\n[train graph {dtype}] VQ indices differing from the oracle's (near-ties): {flips} of {idx_c.numel() + idx_d.numel()}")
    assert dec.shape == (fut, 3, 64, 64) and ref_dec.shape == (2, 3, 64, 64)
    assert rel_err(dec, want[0]) < tol and rel_err(ref_dec, want[1]) < tol
    assert abs(float(commit) - float(want[2])) / float(want[2]) < tol_loss
    assert abs(float(dyn_commit) - float(want[3])) / float(want[3]) < tol_loss
    assert torch.equal(rec.sample, dec) and torch.equal(rec.ref_sample, ref_dec) and float(rec.commit_loss) == float(commit)
    if dtype == torch.bfloat16:      # training arithmetic is fp32 storage / TF32; bf16 is refused, not silently promoted
        with pytest.raises(NotImplementedError, match="fp32"):
            mine(sample=sample.to(cuda), dyn_sample=dyn.to(cuda), segment_len=fut)


@pytest.mark.gpu
@pytest.mark.usefixtures("deterministic")
def test_training_graph_backward_vs_oracle(cuda):
    """Row f3 (backward half): loss.backward() through CompressiveVQModel.forward on the sm_100a kernels against torch autograd
    over the oracle's forward_train (pinned, values and gradients, to the reference's own forward).  The loss is the
    reconstruction + commitment part of train_tokenizer.py:700-712 (MSE of both reconstructions + both commit losses); the
    VQ choices of the product's forward are handed to the oracle (see the forward test).  EVERY parameter's gradient is
    compared: norm-wise relative error per tensor, TF32 tolerance."""
    import torch.nn.functional as F
    from oracle.vq_model_ref import TINY_CFG
    z = np.load(os.path.join(ROOT, "tests", "golden", "tokenizer_refglue.npz"))
    ref, mine = _pair(TINY_CFG, cuda, torch.float32)
    ref.eval(); mine.eval()          # the cross-attention dropouts (conditional_vae.py:24-25, p = 0.1) are random: identity here, tested apart
    px = torch.from_numpy(z["tiny_pixels"])
    fut = px.shape[1] - 2
    sample, dyn = px[0, :2].contiguous(), px[0, 2:].contiguous()
    out = mine(sample=sample.to(cuda), dyn_sample=dyn.to(cuda), segment_len=fut, return_loss=True)
    assert out.sample.requires_grad and out.commit_loss.requires_grad
    loss = F.mse_loss(out.sample, dyn.to(cuda)) + F.mse_loss(out.ref_sample, sample.to(cuda)) + out.commit_loss + 0.5 * out.dyn_commit_loss
    # train_tokenizer.py:706-707 (grad_layer_wrt_loss): gradient of ONE loss term w.r.t. the last decoder layer, graph retained.
    # The sweep stops at the earliest op touching that layer -- the last op of the forward -- so it costs a handful of launches.
    from ivideogpt_b200 import _lib
    n0 = _lib.launch_count()
    probe = torch.autograd.grad(F.mse_loss(out.sample, dyn.to(cuda)), mine.cond_decoder.conv_out.weight, retain_graph=True)[0]
    probe_launches = _lib.launch_count() - n0
    n0 = _lib.launch_count()
    loss.backward()
    full_launches = _lib.launch_count() - n0
    print(f"\n[train graph backward] launches: last-layer probe {probe_launches}, full backward {full_launches}")
    assert probe_launches < 40 and full_launches > 500
    idx = {k: v.cpu() for k, v in mine._last_train_graph.vq_indices.items()}
    # float64 oracle (the product computes in TF32: compare against the exact gradient, not another rounded one)
    ref = ref.double()
    sample64, dyn64 = sample.double(), dyn.double()
    want = ref.forward_train(sample64, dyn64, fut, idx_ctx=idx["ctx"], idx_dyn=idx["dyn"])
    loss_ref = F.mse_loss(want[0], dyn64) + F.mse_loss(want[1], sample64) + want[2] + 0.5 * want[3]
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) / float(loss_ref) < 5e-3
    got = dict(mine.named_parameters())
    worst, missing = [], []
    scale = max(float(p.grad.norm()) for p in ref.parameters() if p.grad is not None)
    for name, p in ref.named_parameters():
        if p.grad is None:
            assert got[name].grad is None or not got[name].grad.any(), name
            continue
        if got[name].grad is None:
            missing.append(name)
            continue
        g_ref, g = p.grad.double(), got[name].grad.double().cpu()
        err = float((g - g_ref).norm() / (g_ref.norm() + 1e-4 * scale))
        worst.append((err, name))
    worst.sort(reverse=True)
    print("\n[train graph backward] parameters compared:", len(worst), " worst:", [(round(e, 4), n) for e, n in worst[:5]])
    if os.environ.get("IVGPT_TEST_VERBOSE"):
        for e, n in worst:
            print(f"   {e:9.5f}  {n}  |g_ref|={float(dict(ref.named_parameters())[n].grad.norm()):.3e}")
    assert not missing, missing[:8]
    assert worst[0][0] < 1e-2, worst[:8]          # measured: 5.6e-3 worst of 316 tensors
    want2 = ref.forward_train(sample64, dyn64, fut, idx_ctx=idx["ctx"], idx_dyn=idx["dyn"])
    probe_ref = torch.autograd.grad(F.mse_loss(want2[0], dyn64), ref.cond_decoder.conv_out.weight)[0]
    assert rel_err(probe, probe_ref) < 5e-3


@pytest.mark.gpu
def test_tokenizer_optimizer_steps_reduce_the_loss(cuda):
    """The reconstruction part of train_tokenizer.py's generator step, end to end on the sm_100a kernels: forward in train mode
    (cross-attention dropouts active) -> loss.backward() -> clip_grad_norm_ -> FusedAdamW.step(), repeated: the loss falls,
    every parameter that has a gradient moves, and the packed weight copies follow the updates (version counters)."""
    import torch.nn.functional as F
    from ivideogpt_b200.optim import FusedAdamW
    from oracle.vq_model_ref import TINY_CFG
    z = np.load(os.path.join(ROOT, "tests", "golden", "tokenizer_refglue.npz"))
    _, mine = _pair(TINY_CFG, cuda, torch.float32)
    mine.train()
    px = torch.from_numpy(z["tiny_pixels"]).to(cuda)
    fut = px.shape[1] - 2
    sample, dyn = px[0, :2].contiguous(), px[0, 2:].contiguous()
    opt = FusedAdamW(mine.parameters(), lr=2e-4, betas=(0.9, 0.99), weight_decay=0.0)
    before = {n: p.detach().clone() for n, p in mine.named_parameters()}
    torch.manual_seed(0)
    losses = []
    for _ in range(6):
        dec, ref_dec, commit, dyn_commit = mine(sample=sample, dyn_sample=dyn, return_dict=False, return_loss=True, segment_len=fut)
        loss = F.mse_loss(dec, dyn) + F.mse_loss(ref_dec, sample) + 0.25 * (commit + dyn_commit)
        loss.backward()
        with_grad = {n for n, p in mine.named_parameters() if p.grad is not None}
        torch.nn.utils.clip_grad_norm_(mine.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(float(loss))
    print("\n[tokenizer training] losses:", [round(l, 4) for l in losses])
    assert losses[-1] < 0.8 * losses[0], losses
    moved = [n for n, p in mine.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert set(moved) == with_grad and len(with_grad) >= 316, sorted(with_grad ^ set(moved))[:8]
    with torch.no_grad():                      # the evaluation forward sees the updated weights (no stale packed copies)
        mine.eval()
        dec_eval = mine(sample=sample, dyn_sample=dyn, segment_len=fut).sample
    assert float(F.mse_loss(dec_eval, dyn)) < float(F.mse_loss(torch.zeros_like(dyn), dyn))


def test_packed_weight_cache_drops_stale_versions():
    """Training bumps every parameter version every step; the kernel-layout cache must follow the update (new copy) and must
    not keep the copies of older versions (it would grow by a full set of packed weights per optimizer step)."""
    import torch.nn as nn
    from ivideogpt_b200.vq_model.plan import PackedWeights
    pw = PackedWeights()
    conv, sc, lin = nn.Conv2d(32, 64, 3, padding=1), nn.Conv2d(32, 64, 1), nn.Linear(64, 64)
    sizes = []
    for step in range(4):
        w, b = pw.conv3(conv, torch.float32, shortcut=sc)
        wd = pw.conv3_dgrad(conv)
        wl, bl = pw.linear(lin.weight, lin.bias, torch.float32)
        wt = pw.linear_t(lin.weight)
        assert w.shape == (64, 9 * 32 + 32) and wd.shape == (32, 9 * 64) and wt.shape == (64, 64)
        # the packed copies reflect the CURRENT values (tf32-rounded)
        assert torch.allclose(w[:, : 9 * 32], conv.weight.detach().permute(0, 2, 3, 1).reshape(64, -1), rtol=1e-3, atol=1e-6)
        assert torch.allclose(wd.view(32, 3, 3, 64)[:, 0, 0], conv.weight.detach()[:, :, 2, 2].t(), rtol=1e-3, atol=1e-6)
        assert torch.allclose(wt, lin.weight.detach().t(), rtol=1e-3, atol=1e-6)
        assert pw.conv3(conv, torch.float32, shortcut=sc)[0] is w          # unchanged parameters: cached
        sizes.append(len(pw._cache))
        with torch.no_grad():                                               # an optimizer step: in-place update, version bump
            for p in list(conv.parameters()) + list(sc.parameters()) + list(lin.parameters()):
                p.add_(0.01)
    assert len(set(sizes)) == 1, sizes


def test_training_forward_refuses_cpu_and_bf16():
    """No fallback: the training graph on CPU tensors raises (there is no torch / CPU path behind it), and a bf16 compute dtype is
    refused for training instead of being silently promoted."""
    from oracle.vq_model_ref import TINY_CFG
    _, mine = _pair(TINY_CFG)
    mine.train()
    s, d = torch.rand(2, 3, 64, 64), torch.rand(2, 3, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mine(sample=s, dyn_sample=d, segment_len=2, return_loss=True)
    from ivideogpt_b200.vq_model.plan import TokenizerPlan
    from ivideogpt_b200.vq_model.train_plan import TokenizerTrainGraph
    with pytest.raises(NotImplementedError, match="fp32"):
        TokenizerTrainGraph(mine, TokenizerPlan(32, torch.bfloat16))


def test_wgrad_index_algebra_of_the_zero_framed_layout():
    """The algebra behind TokenizerTrainGraph._wgrad_conv3, on the CPU in float64: with dY and X written channel-major with a zero
    frame around every image (row pitch Wp), the 3x3 / stride-1 / pad-1 weight gradient of tap (a, b) is the plain product
    dY^T . shift(X^T, (a-1)*Wp + (b-1))^T over the padded pixel axis -- no masking at the image borders, and columns shifted in
    from outside a slice of whole images may be read as zeros (what the TMA does for out-of-range box columns)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    N, H, W, Ci, Co = 3, 6, 10, 4, 5
    x, dy = torch.randn(N, H, W, Ci, generator=g).double(), torch.randn(N, H, W, Co, generator=g).double()
    Wp = (W + 2 + 3) // 4 * 4
    simg = ((H + 2) * Wp + 4 + 63) // 64 * 64

    def framed(t):                                     # [N,H,W,C] -> [C, N*simg]
        p = F.pad(t.permute(3, 0, 1, 2), (1, Wp - W - 1, 1, 1)).reshape(t.shape[-1], N, -1)
        return F.pad(p, (0, simg - p.shape[-1])).reshape(t.shape[-1], N * simg)
    dyT, xT = framed(dy), framed(x)
    cols = F.unfold(x.permute(0, 3, 1, 2), 3, padding=1).view(N, Ci, 9, H * W)
    want = torch.einsum("npo,nctp->otc", dy.view(N, H * W, Co), cols)                 # [Co, tap, Ci]
    for ks in (1, N):                                  # one slice, or one slice per image (out-of-slice columns read as zeros)
        Kc = (N // ks) * simg
        got = torch.zeros(Co, 9, Ci, dtype=torch.float64)
        for a in range(3):
            for b in range(3):
                s = (a - 1) * Wp + (b - 1)
                for k in range(ks):
                    A = dyT[:, k * Kc:(k + 1) * Kc]
                    Bm = torch.zeros(Ci, Kc, dtype=torch.float64)
                    lo, hi = max(0, -s), min(Kc, Kc - s)
                    Bm[:, lo:hi] = xT[:, k * Kc + lo + s: k * Kc + hi + s]
                    got[:, 3 * a + b] += A @ Bm.t()
        assert torch.allclose(got, want, rtol=1e-12, atol=1e-12), ks
