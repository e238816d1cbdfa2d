#!/usr/bin/env python
"""bench.py -- predicted frames/sec of the iVideoGPT next-frame-prediction hot path on B200.

A "step" is one pass of the hot path over one batch of synthetic clips (the body of reference
inference/predict.py:53-73): tokenize the context frames -> autoregressive rollout of the future-frame tokens
(top-k 100 sampling, predict.py:58-63) -> detokenize all frames.

    python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1: independent replicas, weak scaling)
    python bench.py --impl reference ...                      (the reference algorithm's CPU path: oracle + HF Llama)

Prints ONE JSON line (rank 0).  `value` = whole-job predicted frames/s with inputs resident in HBM;
`e2e` = the same metric through the reference-facing Python API exactly as predict.py drives it
(tokenize(all 16 frames) / generate / detokenize) with pinned-host pixels copied H2D and the frames copied D2H
inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (tokenizer config, llama config, resolution, per-GPU batch)
    "cfg64": ("ctx_vae64", "llama_138m", 64, 64),
    "cfg256": ("ctx_vae256", "llama_138m", 256, 16),
    "cfg64-medium": ("ctx_vae64", "llama_436m", 64, 32),
    "tiny": (None, None, 64, 2),
    # train_gpt.py step (BASELINE config 5): frozen tokenizer -> tokens/labels -> Llama fwd+bwd (bf16 compute, fp32
    # master weights) -> gradient all-reduce over ranks -> fused AdamW.  per-device batch 16 (oxe-64-act-free.sh:26)
    "train64": ("ctx_vae64", "llama_138m", 64, 16),
    "train-tiny": (None, None, 64, 2),
    # tokenizer training step (row f3; the reconstruction + commitment part of train_tokenizer.py's generator step):
    # CompressiveVQModel.forward (train mode) -> MSE of both reconstructions + commit losses -> backward -> fused AdamW
    "train-tokenizer64": ("ctx_vae64", "llama_138m", 64, 16),
    "train-tokenizer-tiny": (None, None, 64, 1),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg64", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU clips (default: the workload's)")
    ap.add_argument("--context-length", type=int, default=2)
    ap.add_argument("--segment-length", type=int, default=16)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--greedy", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=1)
    ap.add_argument("--quick", action="store_true", help="profiling aid: exact --warmup, no e2e / cpu legs")
    ap.add_argument("--no-legs", action="store_true", help="skip the nested tf32 / cfg256 / cfg64_medium / train64 / train_tokenizer64 legs")
    return ap.parse_args()


def load_json_cfg(name):
    with open(os.path.join(ROOT, "configs", name + ".json")) as fh:
        return {k: v for k, v in json.load(fh).items() if not k.startswith("_")}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p.get("bf16_tflops_sustained", 1391.4), p.get("hbm_gbs", 6566.7), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# model construction (seeded random weights in the reference state-dict layout; no checkpoints offline)
# ----------------------------------------------------------------------------------------------------
def build_oracle_models(workload):
    import torch
    from oracle.llama_ref import TINY_LLAMA, build_hf_llama, config_path as llama_cfg_path
    from oracle.vq_model_ref import TINY_CFG, RefCompressiveVQModel, seeded_init_
    tok_name, llm_name, _, _ = WORKLOADS[workload]
    tok_cfg = TINY_CFG if tok_name is None else load_json_cfg(tok_name)
    tok_cfg = {k: v for k, v in tok_cfg.items() if k not in ("down_block_types", "up_block_types")}
    ref_tok = seeded_init_(RefCompressiveVQModel(**tok_cfg).eval())
    llm_cfg = dict(TINY_LLAMA, vocab_size=tok_cfg["num_vq_embeddings"] + tok_cfg["num_dyn_embeddings"] + 2) \
        if llm_name is None else llama_cfg_path(llm_name)
    ref_llm = build_hf_llama(llm_cfg)
    return tok_cfg, ref_tok, ref_llm


def build_b200_models(workload, dev, dtype):
    """Same seeded weights as the oracle, loaded through the product classes' state-dict interface."""
    import torch
    from ivideogpt_b200.transformer import B200LlamaForCausalLM
    from ivideogpt_b200.vq_model import CompressiveVQModel
    tok_cfg, ref_tok, ref_llm = build_oracle_models(workload)
    tok = CompressiveVQModel.from_config(tok_cfg)
    tok.load_state_dict(ref_tok.state_dict(), strict=True)
    tok = tok.to(dev).eval().set_compute_dtype(dtype)
    llm = B200LlamaForCausalLM(ref_llm.config)
    llm.load_state_dict(ref_llm.state_dict(), strict=True)
    llm = llm.to(dev).eval().set_compute_dtype(dtype)
    return tok, llm, ref_tok, ref_llm


def host_threads():
    """Threads for the CPU arm.  torch's intra-op pool stops scaling (and then regresses badly) on the small GEMMs
    of a batch-1 rollout well before the 128 hardware threads of the GPU host (the 128-thread run took 322 s for
    one clip); the count actually used is reported as `cores`."""
    return min(os.cpu_count() or 1, int(os.environ.get("IVGPT_CPU_THREADS", "32")))


def synthetic_clips(B, T, res, seed=0):
    import torch
    return torch.rand(B, T, 3, res, res, generator=torch.Generator().manual_seed(seed))


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference algorithm on host cores (oracle tokenizer + genuine HF Llama)
# ----------------------------------------------------------------------------------------------------
def cpu_rollout(ref_tok, ref_llm, clips, ctx, seg, greedy=True):
    import torch
    fut = seg - ctx
    t0 = time.perf_counter()
    with torch.no_grad():
        tokens, _ = ref_tok.tokenize(clips, ctx)          # predict.py:53 tokenizes all frames
        prompt = tokens[:, : ctx * 257]
        out = ref_llm.generate(prompt, do_sample=not greedy, top_k=100, temperature=1.0,
                               max_new_tokens=17 * fut - 1, pad_token_id=50256)
        frames = ref_tok.detokenize(out, ctx).clamp(0.0, 1.0)
    return time.perf_counter() - t0, frames


def cpu_tokenizer_train_step(ref_tok, clips, ctx, seg):
    """The reconstruction + commitment part of train_tokenizer.py's generator step on host cores: oracle forward_train + torch
    autograd + torch AdamW.  Returns seconds for one step over `clips`."""
    import torch
    import torch.nn.functional as F
    ref_tok = ref_tok.train()
    for m in ref_tok.modules():                    # the oracle restates the eval-mode block: keep its dropouts off
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    opt = torch.optim.AdamW(ref_tok.parameters(), lr=1e-4, betas=(0.5, 0.9), weight_decay=0.0)
    B, T = clips.shape[:2]
    s = clips[:, :ctx].reshape(B * ctx, *clips.shape[2:]).contiguous()
    t = clips[:, ctx:].reshape(B * (T - ctx), *clips.shape[2:]).contiguous()
    t0 = time.perf_counter()
    out = ref_tok.forward_train(s, t, seg - ctx)
    (F.mse_loss(out[0], t) + F.mse_loss(out[1], s) + out[2] + out[3]).backward()
    opt.step()
    return time.perf_counter() - t0


def cpu_llm_train_step(ref_tok, ref_llm, clips, ctx):
    """One train_gpt.py step (:776-803) on host cores: oracle tokenizer (frozen) -> HF Llama forward(labels) + backward -> AdamW."""
    import torch
    ref_llm = ref_llm.train()
    opt = torch.optim.AdamW(ref_llm.parameters(), lr=1e-4, weight_decay=0.01)
    t0 = time.perf_counter()
    with torch.no_grad():
        tokens, labels = ref_tok.tokenize(clips, ctx)
    ref_llm(input_ids=tokens, labels=labels).loss.backward()
    opt.step()
    return time.perf_counter() - t0


def run_reference_train(args):
    """--impl reference for the training workloads: the same step on the host cores, bounded to --cpu-clips clips."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = host_threads()
    torch.set_num_threads(cores)
    _, _, res, _ = WORKLOADS[args.workload]
    _, ref_tok, ref_llm = build_oracle_models("tiny" if args.workload.endswith("tiny") else "cfg64")
    ctx, seg = args.context_length, args.segment_length
    clips = synthetic_clips(args.cpu_clips, seg, res)
    tokenizer = args.workload.startswith("train-tokenizer")
    step = (lambda: cpu_tokenizer_train_step(ref_tok, clips, ctx, seg)) if tokenizer else (lambda: cpu_llm_train_step(ref_tok, ref_llm, clips, ctx))
    for _ in range(max(args.warmup, 0) and 1):
        step()
    t = statistics.median([step() for _ in range(max(args.steps, 1))])
    v = args.cpu_clips / t
    what = "oracle forward_train + torch autograd + AdamW" if tokenizer else "oracle tokenizer + HF Llama forward/backward + AdamW"
    sample = f"{args.cpu_clips} clip(s) {res}x{res}x{seg}, fp32, {cores} threads ({what})"
    emit({"impl": "reference", "metric": "tokenizer_train_clips_per_sec" if tokenizer else "train_clips_per_sec", "value": v,
          "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": f"{args.workload}: {sample}", "per_gpu_batch": args.cpu_clips},
          "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
          "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    torch.set_num_threads(cores)
    _, _, res, _ = WORKLOADS[args.workload]
    tok_cfg, ref_tok, ref_llm = build_oracle_models(args.workload)
    ctx, seg = args.context_length, args.segment_length
    clips = synthetic_clips(args.cpu_clips, seg, res)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_rollout(ref_tok, ref_llm, clips, ctx, seg)
    times = [cpu_rollout(ref_tok, ref_llm, clips, ctx, seg)[0] for _ in range(max(args.steps, 1))]
    t = statistics.median(times)
    fps = args.cpu_clips * (seg - ctx) / t
    sample = f"{args.cpu_clips} clip(s) {res}x{res}x{seg}, greedy, fp32, {cores} threads (oracle tokenizer + HF Llama)"
    emit({
        "impl": "reference", "metric": "predicted_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.cpu_clips, res),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args, batch, res):
    ctx, seg = args.context_length, args.segment_length
    return {"workload": f"{args.workload}: {batch} synthetic {res}x{res}x{seg} clips per GPU, context {ctx}, "
                        f"{seg - ctx} predicted frames, {ctx * 257 + 17 * (seg - ctx) - 1} tokens/clip "
                        f"(predict.py defaults; pass --segment-length 17 for the 15-frame variant)",
            "tokenizer": WORKLOADS[args.workload][0], "transformer": WORKLOADS[args.workload][1],
            "per_gpu_batch": batch, "context": ctx, "segment": seg, "predicted": seg - ctx,
            "sampling": "greedy" if args.greedy else "top_k=100,T=1", "parallelism": f"replicas x{args.gpus}",
            "l2_policy": "inputs+weights+activations per step exceed the 126 MB L2 (no explicit flush)"}


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
class Dist:
    """Rank / device / barrier plumbing shared by every leg of one bench process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        import torch.distributed as dist
        if self.world > 1 and dist.is_initialized():
            dist.destroy_process_group()


def measure_inference(args, D, workload, dtype_name, steps, warmup, with_e2e, with_clocks, keep_oracle=False):
    """One inference workload on this process's GPU (all ranks call it together): returns the record of the leg
    (rank 0: full dict; other ranks: None) and, when asked, the oracle models for the CPU baseline."""
    import torch
    from ivideogpt_b200 import _lib
    lib = _lib.load()
    dev = D.dev
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    _, _, res, default_b = WORKLOADS[workload]
    B = args.batch or default_b
    ctx, seg = args.context_length, args.segment_length
    fut = seg - ctx
    max_new = 17 * fut - 1
    tok, llm, ref_tok, ref_llm = build_b200_models(workload, dev, dtype)
    if not keep_oracle:
        ref_tok = ref_llm = None
    clips_host = synthetic_clips(B, seg, res, seed=D.rank).pin_memory()
    clips_dev = clips_host.to(dev)
    gen_kw = dict(do_sample=not args.greedy, temperature=1.0, top_k=100, max_new_tokens=max_new, pad_token_id=50256)

    def step_resident():
        prompt = tok.tokenize_context(clips_dev)                      # prediction-minimal tokenisation
        out = llm.generate(prompt, **gen_kw)
        return tok.detokenize(out, ctx)

    frames_host = torch.empty(B, seg, 3, res, res, dtype=torch.float32).pin_memory()

    def step_e2e():
        px = clips_host.to(dev, non_blocking=True)                    # H2D from pinned memory
        tokens, _ = tok.tokenize(px, ctx)                             # predict.py:53 (all frames)
        out = llm.generate(tokens[:, : ctx * 257], **gen_kw)          # predict.py:54-69
        frames = tok.detokenize(out, ctx).clamp_(0.0, 1.0)            # predict.py:72-73
        frames_host.copy_(frames, non_blocking=True)                  # D2H of the result into pinned memory (stream-ordered:
        return frames_host                                            # the closing CUDA event of the step waits for it)

    def timed(fn, n, profile=False):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        D.barrier()
        n0 = _lib.launch_count()
        if profile:
            lib.ivgpt_profile_enable(1)
        e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_all0.record()
        for i in range(n):
            ev[i][0].record()
            fn()
            ev[i][1].record()
        e_all1.record()
        D.barrier()
        if profile:
            lib.ivgpt_profile_enable(0)
        return D.max_ms(e_all0.elapsed_time(e_all1)), [a.elapsed_time(b) for a, b in ev], _lib.launch_count() - n0

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(D.local) if (D.rank == 0 and with_clocks) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    total_ms, per_step, launches = timed(step_resident, steps, profile=True)
    clocks = sampler.finish() if sampler else None
    # where the step goes (one extra, untimed-for-the-headline pass with events between the three API calls)
    sev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    sev[0].record()
    _prompt = tok.tokenize_context(clips_dev)
    sev[1].record()
    _out = llm.generate(_prompt, **gen_kw)
    sev[2].record()
    tok.detokenize(_out, ctx)
    sev[3].record()
    torch.cuda.synchronize()
    stages = {"tokenize_context_ms": sev[0].elapsed_time(sev[1]), "generate_ms": sev[1].elapsed_time(sev[2]),
              "detokenize_ms": sev[2].elapsed_time(sev[3])}
    # per-launch CUDA-event times of the three kernel families, measured in the timed region above
    prof = {}
    for bucket, name in ((1, "conv"), (0, "gemm"), (2, "mega")):
        ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
        lib.ivgpt_profile_collect(bucket, C.byref(ms), C.byref(fl), C.byref(n))
        prof[name] = (ms.value, fl.value, n.value)
    e2e_steps = max(2, min(steps, 3))
    e2e_ms = float("nan")
    if with_e2e:
        step_e2e()
        e2e_ms, _, _ = timed(step_e2e, e2e_steps)
    rec = None
    if D.rank == 0:
        frames_per_step = D.world * B * fut
        ms_per_step = total_ms / steps
        peak_tf, peak_gbs, peak_src = measured_peaks()
        px_bytes = clips_host.numel() * 4
        rec = {"value": frames_per_step / (ms_per_step / 1e3), "unit": "frames/s", "ms_per_step": ms_per_step,
               "dtype": dtype_name, "per_step_ms": per_step, "stages_ms": stages, "gpu_launches": launches, "clocks": clocks,
               "roofline": build_roofline(args, dtype_name, steps, prof, total_ms, B, ctx, max_new, llm, peak_tf, peak_gbs, peak_src),
               "e2e": {"value": frames_per_step / (e2e_ms / e2e_steps / 1e3) if with_e2e else None, "unit": "frames/s",
                       "h2d_bytes_per_step": px_bytes, "d2h_bytes_per_step": px_bytes,
                       "ms_per_step": e2e_ms / e2e_steps if with_e2e else None,
                       "api": "CompressiveVQModel.tokenize(all frames) -> B200LlamaForCausalLM.generate -> detokenize -> .cpu()"},
               "B": B, "res": res}
    del tok, llm, clips_dev
    torch.cuda.empty_cache()
    return rec, (ref_tok, ref_llm)


def leg(name, fn):
    """Secondary legs never take the headline line down with them."""
    try:
        return fn()
    except Exception as e:        # noqa: BLE001 -- reported in the JSON line
        return {"error": f"{name}: {repr(e)[:300]}"}


def run_b200(args):
    import torch
    D = Dist()
    _, _, res, default_b = WORKLOADS[args.workload]
    B = args.batch or default_b
    ctx, seg = args.context_length, args.segment_length
    fut = seg - ctx
    warmup = args.warmup if args.quick else max(args.warmup, 3)
    main, (ref_tok, ref_llm) = measure_inference(args, D, args.workload, args.dtype, args.steps, warmup, not args.quick, True,
                                                 keep_oracle=(D.world == 1 and not args.no_cpu_baseline and not args.quick))
    out = None
    if D.rank == 0:
        out = {
            "metric": "predicted_frames_per_sec", "value": main["value"], "unit": "frames/s", "n_gpus": D.world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic (U[0,1) pixels, seeded random weights in the reference state-dict layout)",
            "config": workload_config(args, B, res),
            "per_step_ms": main["per_step_ms"], "stages_ms": main["stages_ms"], "gpu_launches": main["gpu_launches"],
            "clocks": main["clocks"], "roofline": main["roofline"], "e2e": main["e2e"],
        }
    legs = not args.quick and not args.no_legs and args.workload == "cfg64" and args.dtype == "bf16" and not args.batch
    if not args.quick:
        r = leg("vq_argmin", lambda: vq_argmin_leg(D.dev, 2 * B * 256, measured_peaks()[1]))
        if out is not None:
            out["vq_argmin"] = r      # BASELINE.json's second metric: GB/s on algorithmic bytes + fp32-pipe fraction
    if legs:
        # The same bench at the parity-grade arithmetic (TF32 tensor cores, fp32 storage: what the reference's predict.py runs
        # on a GPU) and on the other BASELINE.json configurations, as nested records: every driver-run line carries them.
        def short(workload, dtype_name):
            rec, _ = measure_inference(args, D, workload, dtype_name, 2, 3, True, False)
            if rec is None:
                return None
            keep = ("value", "unit", "ms_per_step", "dtype", "stages_ms", "gpu_launches", "B", "res")
            o = {k: rec[k] for k in keep}
            o["e2e"] = {k: rec["e2e"][k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}
            o["roofline"] = {k: rec["roofline"].get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "share_of_step")}
            o["roofline"]["other"] = rec["roofline"].get("other")
            o["steps"], o["warmup"] = 2, 3
            return o
        for key, wl, dt in (("tf32", "cfg64", "tf32"), ("cfg256", "cfg256", "bf16"), ("cfg64_medium", "cfg64-medium", "bf16")):
            r = leg(key, lambda wl=wl, dt=dt: short(wl, dt))
            if out is not None:
                out[key] = r
        r = leg("train64", lambda: train_record(args, D, "train64", 3, 3, cpu_arm=False))
        if out is not None:
            out["train64"] = r
        r = leg("train_tokenizer64", lambda: tok_train_record(args, D, "train-tokenizer64", 3, 3, cpu_arm=False))
        if out is not None:
            out["train_tokenizer64"] = r
    if out is not None and ref_tok is not None:
        cores = host_threads()
        torch.set_num_threads(cores)
        try:
            clips = synthetic_clips(args.cpu_clips, seg, res)
            t_cpu, _ = cpu_rollout(ref_tok, ref_llm, clips, ctx, seg)
            out["cpu_baseline"] = {"value": args.cpu_clips * fut / t_cpu, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": f"{args.cpu_clips} clip {res}x{res}x{seg}, greedy, fp32 (oracle tokenizer + HF Llama), {t_cpu:.1f} s"}
        except Exception as e:     # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": f"failed: {repr(e)[:200]}"}
        if legs:
            # SURVEY 8(d) asks for the CPU path at B = 1 and B = 8; a full B = 8 rollout is minutes of CPU time, so the B = 8
            # sample is bounded: 8 clips, context 2 + 2 predicted frames (33 new tokens per clip)
            def cpu_b8():
                clips8 = synthetic_clips(8, ctx + 2, res)
                t8, _ = cpu_rollout(ref_tok, ref_llm, clips8, ctx, ctx + 2)
                return {"value": 8 * 2 / t8, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": f"8 clips {res}x{res}x{ctx + 2} (2 predicted frames each), greedy, fp32, {t8:.1f} s"}
            out["cpu_baseline_b8"] = leg("cpu_baseline_b8", cpu_b8)
            # the training legs' own CPU arms (1 clip each), last: they update the oracle models' weights
            def cpu_train(kind):
                c1 = synthetic_clips(1, seg, res)
                if kind == "train64":
                    t1 = cpu_llm_train_step(ref_tok, ref_llm, c1, ctx)
                    what = "oracle tokenizer + HF Llama fwd/bwd + AdamW"
                else:
                    t1 = cpu_tokenizer_train_step(ref_tok, c1, ctx, seg)
                    what = "oracle forward_train + torch autograd + AdamW"
                return {"value": 1.0 / t1, "unit": "clips/s", "cores": cores, "kind": "port",
                        "sample": f"1 clip {res}x{res}x{seg}, {what}, fp32, {t1:.1f} s"}
            for kind in ("train64", "train_tokenizer64"):
                if isinstance(out.get(kind), dict) and "error" not in out[kind]:
                    out[kind]["cpu_baseline"] = leg(kind + ".cpu_baseline", lambda kind=kind: cpu_train(kind))
    if out is not None:
        emit(out)
    D.close()


def vq_argmin_leg(dev, N, peak_gbs, K=8192, D=64, iters=10):
    """VQ-argmin kernel alone (ivgpt_vq_argmin, compressive_vq_model.py:199): CUDA events around the call, L2 flushed between
    iterations; algorithmic GB/s and fp32 TFLOP/s (the kernel is fp32-FMA bound)."""
    import torch
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(7)
    z = torch.randn(N, D, generator=g).to(dev)
    e = torch.randn(K, D, generator=g).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.vq_argmin(z, e)
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.vq_argmin(z, e); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    byt = 4.0 * N * D + 4.0 * K * D + 8.0 * N
    out = {"N": N, "K": K, "D": D, "ms": t * 1e3, "algorithmic_GBps": byt / t / 1e9, "frac_of_hbm_peak": byt / t / 1e9 / peak_gbs,
           "fp32_TFLOPs": 2.0 * N * K * D / t / 1e12, "l2_policy": "256 MB flush between iterations"}
    # the kernel is bound by the fp32 FMA pipe: its fraction of the MEASURED FFMA peak of this pool's B200s
    # (tools/probes/mma_probe.cu: 16 independent FFMA chains per thread, 2 x 1024 threads per SM; committed output)
    probe = os.path.join(ROOT, "profiles", "r02", "mma_issue_and_ffma_probe.json")
    if os.path.exists(probe):
        with open(probe) as fh:
            peak = json.load(fh)["ffma"]["reg_reg_tflops"]
        out.update({"fp32_peak_TFLOPs": peak, "frac_of_fp32_peak": out["fp32_TFLOPs"] / peak,
                    "fp32_peak_source": "profiles/r02/mma_issue_and_ffma_probe.json (FFMA micro-benchmark, 3-register form)"})
    return out


def mega_traffic():
    """`traffic` of the dominant kernel from the committed `ncu --set full` capture (bytes per launch), or None."""
    path = os.path.join(ROOT, "profiles", "r02", "ncu_decode_mega_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01", "ncu_decode_mega_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return None


def build_roofline(args, dtype_name, nsteps, prof, total_ms, B, ctx, max_new, llm, peak_tf, peak_gbs, peak_src):
    """Roofline of the DOMINANT kernel (largest summed device time in the timed region).

    decode_mega_kernel is HBM-bound: algorithmic bytes per launch = per decode step the bf16 block
    + lm_head weights once, plus the K and V rows of every earlier position of every clip and layer:
        steps * 2 * (block + lm_head params)  +  sum_{pos} B * pos * layers * 2 * hidden * 2 bytes.
    gemm_tc_kernel is tensor-pipe bound: algorithmic FLOPs 2*M*N*K per launch (causal tiles excluded)."""
    fam = {}
    conv_ms, conv_fl, conv_n = prof["conv"]
    gemm_ms, gemm_fl, gemm_n = prof["gemm"]
    mega_ms, mega_steps, mega_n = prof["mega"]
    dense_peak = peak_tf if dtype_name == "bf16" else peak_tf / 2.0    # tf32 runs at half the bf16 rate
    tsrc = f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src})" + ("" if dtype_name == "bf16" else " / 2 for tf32")
    for name, ms, fl, n, label in (("conv", conv_ms, conv_fl, conv_n, "implicit-GEMM 3x3 conv"),
                                   ("gemm", gemm_ms, gemm_fl, gemm_n, "plain/batched GEMM")):
        ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        fam[name] = {"bound": "tensor", "kernel": f"gemm_tc_kernel<{dtype_name}> ({label})", "achieved": ach,
                     "peak": dense_peak, "unit": "TFLOP/s", "frac": ach / dense_peak, "peak_source": tsrc,
                     "launches_per_step": n / nsteps, "kernel_ms_per_step": ms / nsteps,
                     "share_of_step": ms / total_ms, "traffic": None}
    if mega_n > 0 and mega_ms > 0:
        w = llm.b200_engine().w
        wbytes = 2.0 * (w.layers_n * (4 * w.hidden * w.hidden + 3 * w.hidden * w.inter) + w.vocab * w.hidden)
        L0 = ctx * 257                                   # prompt length; decode step i feeds position L0 + i
        steps = max_new - 1
        kv = sum(B * (L0 + i) * w.layers_n * 2 * w.hidden * 2.0 for i in range(steps))
        per_launch = steps * wbytes + kv
        ach = per_launch * mega_n / (mega_ms * 1e-3) / 1e9
        fam["mega"] = {"bound": "hbm", "kernel": "decode_mega_kernel (persistent decode rollout, one launch per step of the bench)",
                       "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                       "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                       "launches_per_step": mega_n / nsteps, "kernel_ms_per_step": mega_ms / nsteps,
                       "share_of_step": mega_ms / total_ms, "algorithmic_bytes_per_launch": per_launch,
                       "decode_steps_per_launch": steps, "traffic": None}
        cap = mega_traffic()
        if cap:   # ncu --set full DRAM bytes of a shorter launch of the same kernel, scaled by its measured ratio to that launch's algorithmic bytes
            fam["mega"]["traffic"] = cap["ratio_to_algorithmic"] * per_launch
            fam["mega"]["traffic_source"] = (f"{cap.get('source', 'profiles/r01/ncu_decode_mega_traffic.json')}: {cap['dram_bytes'] / 1e9:.2f} GB DRAM over a "
                                             f"{cap['decode_steps']}-step launch = {cap['ratio_to_algorithmic']:.3f} x algorithmic, scaled to {steps} steps")
    dom = max(fam, key=lambda k: fam[k]["kernel_ms_per_step"])
    roofline = dict(fam[dom])
    roofline["other"] = {k: {kk: v[kk] for kk in ("bound", "achieved", "unit", "frac", "kernel_ms_per_step", "share_of_step")}
                         for k, v in fam.items() if k != dom}
    return roofline


def train_record(args, D, workload, steps, warmup, cpu_arm=True):
    """One training step of reference train_gpt.py:766-804 per bench step; metric = clips/s.
    All ranks call it together; rank 0 returns the record."""
    import torch
    import torch.distributed as dist
    from ivideogpt_b200 import _lib
    from ivideogpt_b200.optim import FusedAdamW
    world, rank, dev = D.world, D.rank, D.dev
    base = "tiny" if workload == "train-tiny" else "cfg64"
    _, _, res, default_b = WORKLOADS[workload]
    B = args.batch or default_b
    ctx, seg = args.context_length, args.segment_length
    tok, llm, ref_tok, ref_llm = build_b200_models(base, dev, torch.bfloat16)
    tok.set_compute_dtype(torch.float32)          # train_gpt.py runs the frozen tokenizer in fp32 (TF32 convs)
    llm.train()
    p_drop = float(os.environ.get("IVGPT_TRAIN_ATTN_DROPOUT", "0.0"))    # 0.1 = as scripted (oxe-64-act-free.sh:31); 0 = parity config
    llm.config.attention_dropout = p_drop
    params = [p for p in llm.parameters()]
    decay = [p for p in params if p.dim() >= 2]
    no_decay = [p for p in params if p.dim() < 2]
    # train_gpt.py:640-646: AdamW, weight decay on the matrices only; the exchange is a SUM, the mean is the optimizer's scale
    opt = FusedAdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}], lr=1e-4,
                     betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0 / world)
    clips = synthetic_clips(B, seg, res, seed=rank).to(dev)

    overlap = os.environ.get("IVGPT_TRAIN_OVERLAP", "1") == "1"
    reducer = None
    if world > 1 and overlap:
        # gradient exchange launched bucket by bucket from inside the backward (sum; the mean is AdamW's gscale = 1/world)
        from ivideogpt_b200.grad_reduce import BucketedGradReducer
        reducer = llm.b200_grad_reducer = BucketedGradReducer()

    def train_step():
        with torch.no_grad():
            tokens, labels = tok.tokenize(clips, ctx)                     # train_gpt.py:776-779
        loss = llm(input_ids=tokens, labels=labels).loss                  # :792 (+ overlapped all-reduce when world > 1)
        loss.backward()                                                   # :798
        if world > 1 and not overlap:                                     # one flat all-reduce after the backward (round-1 v1)
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            off = 0
            for p in params:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        opt.step()                                                        # :803 (fused AdamW; bumps every parameter's version:
        opt.zero_grad(set_to_none=True)                                   #  the packed weight copies are rebuilt next step)
        return loss.detach()                                              # :794 gather(loss) deferred: read after the timed region

    for _ in range(max(warmup, 3)):
        train_step()
    sampler = ClockSampler(D.local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    D.barrier()
    n0 = _lib.launch_count()
    b0 = reducer.bytes_reduced if reducer is not None else 0
    k0 = reducer.buckets_launched if reducer is not None else 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    losses = [train_step() for _ in range(steps)]
    b.record()
    D.barrier()
    ms_step = D.max_ms(a.elapsed_time(b)) / steps
    clocks = sampler.finish() if sampler else None
    rec = None
    if rank == 0:
        tokens_per_clip = ctx * 257 + 17 * (seg - ctx) - 1
        flops_clip = 6.0 * 125.8e6 * tokens_per_clip + 31e9            # ~567 + 31 GFLOP per clip (fwd+bwd)
        peak_tf, _, src = measured_peaks()
        ach = B * flops_clip / (ms_step * 1e-3) / 1e12
        if world == 1:
            exch = {"algorithm": "none (1 GPU)"}
        elif reducer is not None:
            exch = {"algorithm": "bucketed NCCL all-reduce (SUM), one bucket per layer launched from inside the backward, overlapped",
                    "buckets_per_step": (reducer.buckets_launched - k0) / steps,
                    "bytes_per_step": (reducer.bytes_reduced - b0) / steps,
                    "bucket_bytes": {"lm_head+final_norm": 4 * (16386 * 768 + 768), "layer": 4 * (12 * 768 * 768 + 2 * 768),
                                     "embedding": 4 * 16386 * 768}}
        else:
            exch = {"algorithm": "one flat NCCL all-reduce after the backward"}
        rec = {
            "metric": "train_clips_per_sec", "value": world * B / (ms_step / 1e3), "unit": "clips/s", "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{workload}: train_gpt.py step, {B} clips/GPU of {res}x{res}x{seg}, frozen fp32 "
                                   f"tokenizer -> Llama fwd+bwd bf16 -> all-reduce -> AdamW", "per_gpu_batch": B,
                       "parallelism": f"dp{world}", "attention_dropout": p_drop},
            "grad_exchange": exch,
            "loss_first_last": [float(losses[0]), float(losses[-1])], "gpu_launches": _lib.launch_count() - n0,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "training step (all tcgen05 GEMMs)", "achieved": ach, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": ach / peak_tf, "peak_source": f"bf16_tflops_sustained ({src})",
                         "traffic": None, "note": "model-FLOPs utilisation (MFU) of the transformer fwd+bwd only, per GPU"},
        }
        # the same step on the host cores (bounded: 1 clip): oracle tokenizer + HF Llama forward/backward + torch AdamW
        # (nested legs of a default run get it at the very end of the run instead: CPU work between GPU legs perturbs them)
        if cpu_arm:
            try:
                cores = host_threads()
                torch.set_num_threads(cores)
                t_cpu = cpu_llm_train_step(ref_tok, ref_llm, synthetic_clips(1, seg, res), ctx)
                rec["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "clips/s", "cores": cores, "kind": "port",
                                       "sample": f"1 clip {res}x{res}x{seg}, oracle tokenizer + HF Llama fwd/bwd + AdamW, fp32, {t_cpu:.1f} s"}
            except Exception as e:     # noqa: BLE001
                rec["cpu_baseline"] = {"value": None, "unit": "clips/s", "kind": "port", "sample": f"failed: {repr(e)[:200]}"}
    del tok, llm, opt, params
    torch.cuda.empty_cache()
    return rec


def tok_train_record(args, D, workload, steps, warmup, cpu_arm=True):
    """One tokenizer training step per bench step (reference train_tokenizer.py:620-740 without the LPIPS / GAN terms, which are
    out of scope): forward in train mode, loss = MSE(dec, target) + MSE(ref_dec, context) + commit + dyn_commit, backward on
    the sm_100a kernels, gradient all-reduce (N > 1), clip_grad_norm_, fused AdamW.  metric = clips/s."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from ivideogpt_b200 import _lib
    from ivideogpt_b200.grad_reduce import allreduce_grads_flat
    from ivideogpt_b200.optim import FusedAdamW
    world, rank, dev = D.world, D.rank, D.dev
    base = "tiny" if workload.endswith("tiny") else "cfg64"
    _, _, res, default_b = WORKLOADS[workload]
    B = args.batch or default_b
    ctx, seg = args.context_length, args.segment_length
    tok, llm, ref_tok, _ = build_b200_models(base, dev, torch.float32)
    del llm
    tok.train()
    params = [p for p in tok.parameters()]
    opt = FusedAdamW(params, lr=1e-4, betas=(0.5, 0.9), weight_decay=0.0, grad_scale=1.0 / world)   # train_tokenizer.py:466-472
    clips = synthetic_clips(B, seg, res, seed=rank).to(dev)
    sample = clips[:, :ctx].reshape(B * ctx, 3, res, res).contiguous()
    target = clips[:, ctx:].reshape(B * (seg - ctx), 3, res, res).contiguous()

    def step():
        dec, ref_dec, commit, dyn_commit = tok(sample=sample, dyn_sample=target, return_dict=False, return_loss=True,
                                               segment_len=seg - ctx)
        loss = F.mse_loss(dec, target) + F.mse_loss(ref_dec, sample) + commit + dyn_commit
        loss.backward()
        allreduce_grads_flat(params)              # SUM over ranks (no-op at N = 1); the mean is AdamW's grad_scale = 1/world
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.detach()

    for _ in range(max(warmup, 3)):
        step()
    sampler = ClockSampler(D.local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    D.barrier()
    n0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    losses = [step() for _ in range(steps)]
    b.record()
    D.barrier()
    ms_step = D.max_ms(a.elapsed_time(b)) / steps
    clocks = sampler.finish() if sampler else None
    rec = None
    if rank == 0:
        mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        rec = {
            "metric": "tokenizer_train_clips_per_sec", "value": world * B / (ms_step / 1e3), "unit": "clips/s", "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": f"{workload}: CompressiveVQModel.forward (train mode) + backward + AdamW, {B} clips/GPU of "
                                   f"{res}x{res}x{seg} (context {ctx}), loss = MSE + commit (no LPIPS / GAN terms)",
                       "per_gpu_batch": B, "parallelism": f"dp{world}"},
            "loss_first_last": [float(losses[0]), float(losses[-1])], "gpu_launches": _lib.launch_count() - n0,
            "peak_memory_gib": mem, "clocks": clocks,
        }
        # the same step on the host cores: the oracle's differentiable forward_train + torch autograd + torch AdamW, 1 clip
        if cpu_arm:
            try:
                cores = host_threads()
                torch.set_num_threads(cores)
                t_cpu = cpu_tokenizer_train_step(ref_tok, synthetic_clips(1, seg, res), ctx, seg)
                rec["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "clips/s", "cores": cores, "kind": "port",
                                       "sample": f"1 clip {res}x{res}x{seg}, oracle forward_train + torch autograd + AdamW, fp32, {t_cpu:.1f} s"}
            except Exception as e:     # noqa: BLE001
                rec["cpu_baseline"] = {"value": None, "unit": "clips/s", "kind": "port", "sample": f"failed: {repr(e)[:200]}"}
    del tok, opt, params
    torch.cuda.empty_cache()
    return rec


def run_train(args):
    D = Dist()
    if args.workload.startswith("train-tokenizer"):
        rec = tok_train_record(args, D, args.workload, args.steps, args.warmup)
        if rec is not None:
            emit(rec)
        D.close()
        return
    rec = train_record(args, D, args.workload, args.steps, args.warmup)
    if rec is not None:
        emit(rec)
    D.close()


def emit(rec):
    """The one JSON line, on the process's real stdout (see _quiet_stdout)."""
    os.write(_REAL_STDOUT, (json.dumps(rec) + "\n").encode())


def _quiet_stdout():
    """NCCL (NCCL_DEBUG=VERSION on some boxes) and library banners write to fd 1; the contract is ONE line on stdout, so fd 1
    is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse()
    _quiet_stdout()
    if a.impl == "reference":
        (run_reference_train if a.workload.startswith("train") else run_reference)(a)
    elif a.workload.startswith("train"):
        run_train(a)
    else:
        run_b200(a)
