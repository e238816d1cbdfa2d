"""The drop-in claim, executed: the reference's OWN driver code runs against this package with no edit.

  * `inference/predict.py` is run as a subprocess, byte-identical (sha256 checked), against a checkpoint directory written
    by save_pretrained from seeded weights; PYTHONPATH holds the repo root (where `ivideogpt/` is the alias package of
    ivideogpt_b200) and a stand-in for the one missing third-party import (`imageio`, tests/shims/).
  * the training loop body of `train_gpt.py` (the `for step, batch in enumerate(active_dataloader)` block up to
    `lr_scheduler.step()`, :766-804) is extracted from the reference file at run time and exec'd, byte-identical, over a
    minimal stand-in for the `accelerate.Accelerator` object (accelerate is not installed; SURVEY App. D.7 ii) with the
    model / tokenizer / optimizer built the way train_gpt.py builds them (AutoModelForCausalLM.from_config -> our class,
    torch.optim.AdamW, get_scheduler).

The reference files are looked up in $IVGPT_REFERENCE_ROOT, /root/reference (build container) or baseline/_ref (unmodified
copies staged by tools/stage_reference_drivers.sh; git-ignored, travels to the GPU box).  Without them the tests skip.
"""
import hashlib
import os
import subprocess
import sys
import textwrap
from types import SimpleNamespace

import pytest
import torch

from helpers import ROOT

SHA = {"inference/predict.py": "d3baf86a052693e2e9d712241503325e5d8d48a25afb7f6db6b16ee78eb443df",
       "inference/utils.py": "ec533156eabed4c2aae48b3702e7072c7bb3506f7cc0fa29168805015948eda2",
       "train_gpt.py": "aee87cd9f748774984ee8e2ff0b0dc8f97865cfac49e14329d0f2bd94a486e45",
       "train_tokenizer.py": "f46f2e455af3bbe21202bba690ac4df9eced014874f44108d7da3389047404e2"}


def _ref_root():
    for cand in (os.environ.get("IVGPT_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "inference", "predict.py")) and os.path.isfile(os.path.join(cand, "train_gpt.py")):
            return cand
    return None


def _check_unmodified(root, rel):
    with open(os.path.join(root, rel), "rb") as fh:
        assert hashlib.sha256(fh.read()).hexdigest() == SHA[rel], f"{rel} is not the reference's file (sha256 differs)"


def _seeded_checkpoint(tmp_path):
    """<dir>/tokenizer + <dir>/transformer written by the product classes' save_pretrained from seeded weights (tiny
    configuration with the released token geometry: 256 tokens per context frame, 16 per future frame)."""
    from oracle.llama_ref import TINY_LLAMA, build_hf_llama
    from oracle.vq_model_ref import TINY_CFG, RefCompressiveVQModel, seeded_init_
    from ivideogpt_b200.transformer import B200LlamaForCausalLM
    from ivideogpt_b200.vq_model import CompressiveVQModel
    ref_tok = seeded_init_(RefCompressiveVQModel(**TINY_CFG).eval())
    tok = CompressiveVQModel.from_config(TINY_CFG)
    tok.load_state_dict(ref_tok.state_dict(), strict=True)
    tok.save_pretrained(str(tmp_path / "ckpt" / "tokenizer"))
    hf = build_hf_llama(TINY_LLAMA, init_scale=3.0)
    llm = B200LlamaForCausalLM(hf.config)
    llm.load_state_dict(hf.state_dict(), strict=True)
    llm.save_pretrained(str(tmp_path / "ckpt" / "transformer"))
    return str(tmp_path / "ckpt")


@pytest.mark.gpu
def test_reference_predict_py_runs_unmodified(cuda, tmp_path):
    root = _ref_root()
    if root is None:
        pytest.skip("reference driver files not available (run tools/stage_reference_drivers.sh in the build container)")
    _check_unmodified(root, "inference/predict.py")
    _check_unmodified(root, "inference/utils.py")
    ckpt = _seeded_checkpoint(tmp_path)
    out_dir = tmp_path / "out"
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "tests", "shims"), env.get("PYTHONPATH", "")])
    cmd = [sys.executable, os.path.join(root, "inference", "predict.py"), "--pretrained_model_name_or_path", ckpt,
           "--input_path", os.path.join(root, "inference", "samples", "bair_sample.npz"), "--dataset_name", "bair_robot_pushing",
           "--output_path", str(out_dir), "--repeat_times", "2", "--segment_length", "6", "--context_length", "2"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, f"predict.py failed:\nSTDOUT:\n{r.stdout[-2000:]}\nSTDERR:\n{r.stderr[-4000:]}"
    gifs = sorted(p.name for p in out_dir.iterdir())
    assert gifs == ["pred-samples-0.gif", "pred-samples-1.gif"], gifs
    from PIL import Image
    im = Image.open(out_dir / "pred-samples-0.gif")
    assert im.size == (128, 64) and getattr(im, "n_frames", 1) == 6       # ground truth | prediction, 6 frames


@pytest.mark.gpu
def test_reference_train_gpt_loop_body_runs_unmodified(cuda, tmp_path):
    root = _ref_root()
    if root is None:
        pytest.skip("reference driver files not available (run tools/stage_reference_drivers.sh in the build container)")
    _check_unmodified(root, "train_gpt.py")
    with open(os.path.join(root, "train_gpt.py")) as fh:
        lines = fh.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "for step, batch in enumerate(active_dataloader):")
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == "lr_scheduler.step()")
    assert (start + 1, end + 1) == (766, 804), (start + 1, end + 1)           # the block SURVEY / VERDICT cite
    body = textwrap.dedent("\n".join(lines[start:end + 1]))

    # ---- objects built the way train_gpt.py builds them ----
    import ivideogpt.transformer  # noqa: F401  (train_gpt.py:43 -- registers the B200 class with AutoModelForCausalLM)
    from transformers import AutoModelForCausalLM, LlamaConfig, get_scheduler
    from ivideogpt.vq_model import CompressiveVQModel
    from ivideogpt_b200.transformer import B200LlamaForCausalLM
    from oracle.llama_ref import TINY_LLAMA
    from oracle.vq_model_ref import TINY_CFG, RefCompressiveVQModel, seeded_init_
    torch.manual_seed(0)
    ref_tok = seeded_init_(RefCompressiveVQModel(**TINY_CFG).eval())
    tokenizer = CompressiveVQModel.from_config(TINY_CFG)
    tokenizer.load_state_dict(ref_tok.state_dict(), strict=True)
    tokenizer = tokenizer.to(cuda).eval()
    model = AutoModelForCausalLM.from_config(LlamaConfig(**TINY_LLAMA))         # train_gpt.py:597
    assert isinstance(model, B200LlamaForCausalLM)
    model = model.to(cuda)
    optimizer = torch.optim.AdamW(model.parameters(), lr=3e-3, weight_decay=0.01)   # train_gpt.py:640-646
    lr_scheduler = get_scheduler("constant", optimizer=optimizer, num_warmup_steps=0, num_training_steps=10)

    losses = []

    class Accel:      # the handful of accelerate.Accelerator members the loop body touches
        device = cuda
        sync_gradients = True
        is_main_process = True

        def unwrap_model(self, m):
            return m

        def accumulate(self, m):
            import contextlib
            return contextlib.nullcontext()

        def gather(self, t):
            return t

        def backward(self, loss):
            losses.append(float(loss.detach()))
            loss.backward()

        def clip_grad_norm_(self, params, max_norm):
            return torch.nn.utils.clip_grad_norm_(params, max_norm)

    args = SimpleNamespace(action_conditioned=False, reward_prediction=False, action_recon=None, context_length=2,
                           per_device_train_batch_size=2, max_grad_norm=1.0)
    g = torch.Generator().manual_seed(0)
    batch = torch.rand(2, 6, 3, 64, 64, generator=g)
    ns = dict(torch=torch, args=args, accelerator=Accel(), tokenizer=tokenizer, model=model, optimizer=optimizer,
              lr_scheduler=lr_scheduler, active_dataloader=[batch, batch, batch, batch])
    model.train()
    exec(compile(body, os.path.join(root, "train_gpt.py"), "exec"), ns)        # the reference's own loop body, 4 steps
    assert len(losses) == 4 and all(l == l and l < 1e4 for l in losses), losses
    assert float(ns["avg_loss"]) == pytest.approx(losses[-1], rel=1e-6)
    # same batch every step with a healthy learning rate: the loss must go down -- parameters (and the packed kernel-layout
    # copies the engine rebuilds after every torch.optim step) really are being updated
    assert losses[-1] < losses[0] - 0.05, losses


@pytest.mark.gpu
def test_reference_train_tokenizer_generator_step_runs_unmodified(cuda):
    """Row f3: the generator step of the reference's `train_tokenizer.py` (the `for i, batch in enumerate(train_dataloader)` body
    from :583 to the optimizer / scheduler step at :744-745, and its helper `grad_layer_wrt_loss` :64-70) is exec'd UNMODIFIED
    against this package's CompressiveVQModel: `model(sample=, dyn_sample=, return_dict=False, return_loss=True, segment_len=)`
    in train mode, the loss assembled by the reference's own code, `accelerator.backward(loss)` through the sm_100a backward,
    clip_grad_norm_, optimizer.step().  Two regimes: before `disc_start` (reconstruction + perceptual + commit terms) and after
    it, where the reference takes `torch.autograd.grad(..., retain_graph=True)` of two losses w.r.t. the last decoder layer
    (the adaptive GAN weight) before the real backward.  LPIPS and the discriminator are out of scope: small
    frozen torch conv nets stand in for them -- what is exercised is the tokenizer's forward/backward under the reference's
    own control flow."""
    import time
    import torch.nn as nn
    import torch.nn.functional as F
    root = _ref_root()
    if root is None or not os.path.isfile(os.path.join(root, "train_tokenizer.py")):
        pytest.skip("reference driver files not available (run tools/stage_reference_drivers.sh in the build container)")
    _check_unmodified(root, "train_tokenizer.py")
    with open(os.path.join(root, "train_tokenizer.py")) as fh:
        lines = fh.read().split("\n")
    h0 = next(i for i, l in enumerate(lines) if l.startswith("def grad_layer_wrt_loss"))
    helper = "\n".join(lines[h0:h0 + 7])
    start = next(i for i, l in enumerate(lines) if l.strip() == "for i, batch in enumerate(train_dataloader):")
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == "lr_scheduler.step()")
    assert (h0 + 1, start + 1, end + 1) == (64, 583, 745), (h0 + 1, start + 1, end + 1)
    body = textwrap.dedent("\n".join(lines[start:end + 1]))

    from ivideogpt.vq_model import CompressiveVQModel
    from oracle.vq_model_ref import TINY_CFG, RefCompressiveVQModel, seeded_init_
    torch.manual_seed(0)
    ref_tok = seeded_init_(RefCompressiveVQModel(**TINY_CFG).eval())
    model = CompressiveVQModel.from_config(TINY_CFG)
    model.load_state_dict(ref_tok.state_dict(), strict=True)
    model = model.to(cuda).train()
    optimizer = torch.optim.AdamW(model.parameters(), lr=2e-4, betas=(0.5, 0.9), weight_decay=0.0)      # train_tokenizer.py:466-472
    feat = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.SiLU(), nn.Conv2d(8, 8, 3, stride=2, padding=1)).to(cuda).requires_grad_(False)
    discriminator = nn.Sequential(nn.Conv2d(3, 8, 4, stride=2, padding=1), nn.LeakyReLU(0.2), nn.Conv2d(8, 1, 4, stride=2, padding=1)).to(cuda)

    def lpips(a, b, weight=None):                    # stand-in with LPIPS's call signature: per-sample feature distance
        return ((feat(a) - feat(b)) ** 2).mean(dim=(1, 2, 3))

    losses = []

    class Accel:
        device = cuda
        sync_gradients = True
        is_main_process = True
        is_local_main_process = True

        def unwrap_model(self, m):
            return m

        def accumulate(self, m):
            import contextlib
            return contextlib.nullcontext()

        def gather(self, t):
            return t

        def backward(self, loss):
            losses.append(float(loss.detach()))
            loss.backward()

        def clip_grad_norm_(self, params, max_norm):
            return torch.nn.utils.clip_grad_norm_(params, max_norm)

    class Meter:
        def update(self, v):
            pass

    class Sched:
        def step(self):
            pass

    args = SimpleNamespace(model_type="ctx_vqgan", train_batch_size=2, segment_length=6, context_length=2,
                           gradient_accumulation_steps=1, disc_start=100, vae_loss="l2", balanced_loss=False, recon_weight=1.0,
                           perc_weight=1.0, disc_weight=0.1, weighted_gan=False, max_grad_norm=1.0, log_grad_norm_steps=10 ** 9)
    g = torch.Generator().manual_seed(0)
    batch = torch.rand(2, 6, 3, 64, 64, generator=g)
    ns = dict(torch=torch, F=F, time=time, args=args, accelerator=Accel(), model=model, optimizer=optimizer,
              discr_optimizer=torch.optim.AdamW(discriminator.parameters(), lr=1e-4), lr_scheduler=Sched(), lpips=lpips,
              discriminator=discriminator, data_time_m=Meter(), end=time.time(), global_step=0, log_grad_norm=lambda *a: None,
              train_dataloader=[batch] * 12)      # even i are generator steps: 6 optimizer steps
    exec(compile(helper, os.path.join(root, "train_tokenizer.py"), "exec"), ns)
    exec(compile(body, os.path.join(root, "train_tokenizer.py"), "exec"), ns)
    assert len(losses) == 6 and all(l == l and l < 1e4 for l in losses), losses
    # same batch: after the first Adam steps' transient (measured: 2.6 -> 6.9 -> 3.0 -> ...) the tokenizer is learning
    assert losses[-1] < losses[0] and losses[-1] < 0.5 * max(losses), losses
    assert ns["fmap"].shape == (8, 3, 64, 64) and ns["fmap_ref"].shape == (4, 3, 64, 64)
    # after disc_start: generator loss through the discriminator + the adaptive weight probes (retain_graph) + real backward
    ns["global_step"] = 100
    ns["train_dataloader"] = [batch]
    n_before = len(losses)
    exec(compile(body, os.path.join(root, "train_tokenizer.py"), "exec"), ns)
    assert len(losses) == n_before + 1 and losses[-1] == losses[-1]
    aw = float(ns["adaptive_weight"])
    assert 0.0 < aw <= 1e4, aw
    assert float(ns["avg_gan_loss"]) == float(ns["avg_gan_loss"])
