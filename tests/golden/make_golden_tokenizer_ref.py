"""Regenerates tests/golden/tokenizer_refglue.npz by RUNNING THE REFERENCE'S OWN TOKENIZER FILES
(/root/reference/ivideogpt/vq_model/{vae,conditional_vae,compressive_vq_model}.py, unmodified) on top of
oracle/diffusers_stub -- a stand-in for the few diffusers 0.27.0 symbols those files import, built from the oracle's
restatements of the published diffusers blocks (diffusers itself is not installable offline).

What this pins: everything the reference itself wrote on the path -- Encoder / Decoder wiring and feature taps,
CrossAttentionBlock, Conditional{En,De}coder, tokenize / detokenize (incl. cache / return_cache), token serialisation and
labels -- executes as shipped; the oracle (oracle/vq_model_ref.py) must reproduce its outputs.  What stays restated: the
diffusers building blocks inside the stub.  Run from the repo root in the build container:
    python tests/golden/make_golden_tokenizer_ref.py
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_stub"))
REF = "/root/reference/ivideogpt/vq_model"
spec = importlib.util.spec_from_file_location("ref_vq_model", REF + "/__init__.py", submodule_search_locations=[REF])
ref_pkg = importlib.util.module_from_spec(spec)
sys.modules["ref_vq_model"] = ref_pkg
spec.loader.exec_module(ref_pkg)

from oracle.vq_model_ref import TINY_CFG, RefCompressiveVQModel, config_path, seeded_init_  # noqa: E402

torch.set_num_threads(1)


def load_cfg(name):
    with open(config_path(name)) as fh:
        return {k: v for k, v in json.load(fh).items() if not k.startswith("_")}


def run(cfg, frames, res, seed, want_cache):
    ocfg = {k: v for k, v in cfg.items() if k not in ("down_block_types", "up_block_types")}
    n = len(cfg["block_out_channels"])
    oracle = seeded_init_(RefCompressiveVQModel(**ocfg).eval())
    reference = ref_pkg.CompressiveVQModel(**dict(ocfg, down_block_types=("DownEncoderBlock2D",) * n,
                                                  up_block_types=("UpDecoderBlock2D",) * n)).eval()
    assert set(reference.state_dict()) == set(oracle.state_dict()), "state-dict keys differ from the reference module tree"
    reference.load_state_dict(oracle.state_dict(), strict=True)
    ctx = cfg["context_length"]
    px = torch.rand(1, frames, 3, res, res, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        tok, lab = reference.tokenize(px, ctx)
        rec = reference.detokenize(tok, ctx)
        otok, olab = oracle.tokenize(px, ctx)
        orec = oracle.detokenize(otok, ctx)
        out = dict(pixels=px.numpy(), tokens=tok.numpy(), labels=lab.numpy(), recon=rec.numpy().astype(np.float32))
        if want_cache:      # compressive_vq_model.py:253-277: context decoded once, future frames decoded one by one from the cache
            # (the reference's cache holds the context features already repeated per future frame, so it only works for
            #  single-future-frame calls -- the MBRL rollout pattern, mbrl/video_predictor.py:320-321)
            c0 = ctx * 257 - 1
            _, cache = reference.detokenize(tok[:, : c0 + 17], ctx, return_cache=True)
            seq2 = torch.cat([tok[:, :c0], tok[:, c0 + 17: c0 + 34]], dim=1)
            step = reference.detokenize(seq2, ctx, cache=cache)
            out["cached_tokens"] = seq2.numpy()
            out["recon_cached"] = step.numpy().astype(np.float32)
    diff = dict(tokens_equal=bool((tok == otok).all()), labels_equal=bool((lab == olab).all()),
                recon_max_abs_diff=float((rec - orec).abs().max()))
    if want_cache:
        # the TRAINING graph (compressive_vq_model.py:332-369 forward + :290-330 decode): values and gradients of the reference's
        # own forward() against the oracle's forward_train() -- straight-through VQ, both commit losses, both decoders
        fut = frames - ctx
        sample = px[0, :ctx].clone()
        dyn = px[0, ctx:].clone()
        ra = reference(sample=sample, dyn_sample=dyn, segment_len=fut, return_dict=False, return_loss=True)
        oa = oracle.forward_train(sample, dyn, fut)
        diff["train_forward_max_abs_diff"] = max(float((x - y).abs().max()) for x, y in zip(ra, oa))
        for outs, model in ((ra, reference), (oa, oracle)):
            for p in model.parameters():
                p.grad = None
            (sum(((x - 0.5) ** 2).mean() for x in outs[:2]) + outs[2] + outs[3]).backward()
        og = dict(oracle.named_parameters())
        gd = 0.0
        for n_, p in reference.named_parameters():
            if p.grad is not None:
                gd = max(gd, float((p.grad - og[n_].grad).abs().max() / p.grad.abs().max().clamp_min(1e-12)))
        diff["train_grad_max_rel_diff"] = gd
        out["train_dec"] = ra[0].detach().numpy().astype(np.float32)
        out["train_ref_dec"] = ra[1].detach().numpy().astype(np.float32)
        out["train_losses"] = np.array([float(ra[2]), float(ra[3])], dtype=np.float64)
        out["train_grad_norms"] = np.array([float(reference.quant_conv.weight.grad.norm()),
                                            float(reference.cond_decoder.conv_out.weight.grad.norm()),
                                            float(reference.quantize.embedding.weight.grad.norm()),
                                            float(reference.encoder.conv_in.weight.grad.norm())], dtype=np.float64)
    return out, diff


summary = {}
fixture = {}
for name, cfg, frames, res, seed, cache in (("tiny", TINY_CFG, 5, 64, 11, True), ("cfg64", load_cfg("ctx_vae64"), 4, 64, 12, False)):
    out, diff = run(dict(cfg), frames, res, seed, cache)
    summary[name] = diff
    for k, v in out.items():
        fixture[f"{name}_{k}"] = v
if os.environ.get("WITH_256", "1") == "1":          # executed here, not stored (a 256x256 clip is too big for a fixture)
    _, summary["cfg256_not_stored"] = run(load_cfg("ctx_vae256"), 3, 256, 13, False)
print(json.dumps(summary, indent=1))
assert all(d["tokens_equal"] and d["labels_equal"] and d["recon_max_abs_diff"] < 1e-5 for d in summary.values()), summary
assert summary["tiny"]["train_forward_max_abs_diff"] < 1e-5 and summary["tiny"]["train_grad_max_rel_diff"] < 1e-4, summary
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tokenizer_refglue.npz"), summary=np.array(json.dumps(summary)), **fixture)
