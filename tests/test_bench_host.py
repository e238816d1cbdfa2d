"""Host-side logic of bench.py (no GPU): which kernel family the roofline object reports, the algorithmic-byte formula of
the decode megakernel (251.7 MB of weights per step + B * pos * 36 864 B of K/V), the launch-list summariser."""
import os
import subprocess
import sys
from types import SimpleNamespace

import pytest

from helpers import ROOT

sys.path.insert(0, ROOT)


def _llm(hidden=768, inter=3072, layers=12, vocab=16386):
    w = SimpleNamespace(hidden=hidden, inter=inter, layers_n=layers, vocab=vocab)
    return SimpleNamespace(b200_engine=lambda: SimpleNamespace(w=w))


def test_roofline_reports_the_dominant_kernel_and_its_algorithmic_bytes():
    import bench
    args = SimpleNamespace(dtype="bf16", steps=5)
    # (summed ms, summed work, launches) over 5 steps: decode megakernel dominates
    prof = {"conv": (225.0, 2.1e14, 345), "gemm": (105.0, 4.5e13, 600), "mega": (950.0, 5 * 236.0, 5)}
    r = bench.build_roofline(args, "bf16", 5, prof, 1450.0, 64, 2, 237, _llm(), 1349.9, 6530.3, "measured")
    assert r["bound"] == "hbm" and r["kernel"].startswith("decode_mega_kernel") and r["unit"] == "GB/s"
    # 236 steps from a 514-token prompt at B=64: 59.4 GB of weights + 351.6 GB of K/V
    assert abs(r["algorithmic_bytes_per_launch"] - 411.0e9) < 0.5e9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] * 5 / 0.950 / 1e9) < 1e-6
    assert abs(r["frac"] - r["achieved"] / 6530.3) < 1e-12 and abs(r["share_of_step"] - 950.0 / 1450.0) < 1e-12
    assert set(r["other"]) == {"conv", "gemm"} and r["other"]["conv"]["bound"] == "tensor"
    assert abs(r["other"]["conv"]["achieved"] - 2.1e14 / 0.225 / 1e12) < 1e-6
    # without the megakernel (TF32 parity path) the conv family is reported, against half the bf16 peak
    args32 = SimpleNamespace(dtype="tf32", steps=5)
    r2 = bench.build_roofline(args32, "tf32", 5, dict(prof, mega=(0.0, 0.0, 0)), 1450.0, 64, 2, 237, _llm(), 1349.9, 6530.3, "measured")
    assert r2["bound"] == "tensor" and "conv" in r2["kernel"] and abs(r2["peak"] - 1349.9 / 2) < 1e-9


def test_workload_table_matches_baseline_configs():
    import bench
    assert bench.WORKLOADS["cfg64"][2:] == (64, 64) and bench.WORKLOADS["cfg256"][2:] == (256, 16)
    assert bench.WORKLOADS["cfg64-medium"][1] == "llama_436m" and bench.WORKLOADS["train64"][3] == 16
    a = SimpleNamespace(workload="cfg64", context_length=2, segment_length=16, greedy=False, gpus=1)
    cfg = bench.workload_config(a, 64, 64)
    assert cfg["predicted"] == 14 and "751 tokens/clip" in cfg["workload"] and "model" not in cfg


def test_launch_list_summariser(tmp_path):
    csv = tmp_path / "l.csv"
    csv.write_text('==PROF== x\n"ID","Kernel Name","Metric Name","Metric Unit","Metric Value"\n'
                   '"0","void ivg::k1<int>(int)","gpu__time_duration.sum","us","10.0"\n'
                   '"1","void ivg::k1<int>(int)","gpu__time_duration.sum","us","30.0"\n'
                   '"2","ivg::k2(float)","gpu__time_duration.sum","ns","60000"\n')
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarise_launches.py"), str(csv)],
                         capture_output=True, text=True, check=True).stdout
    assert "launches 3" in out and "60.00%" in out and "2 x" in out


@pytest.mark.parametrize("workload,metric", [("train-tokenizer-tiny", "tokenizer_train_clips_per_sec"), ("train-tiny", "train_clips_per_sec")])
def test_reference_arm_of_the_training_workloads(workload, metric):
    """`bench.py --impl reference --workload train*`: the same training step on the host cores (oracle tokenizer / HF Llama +
    torch autograd + torch AdamW), ONE JSON line on stdout with the reference-arm keys of the contract."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1",
                          "--warmup", "0", "--cpu-clips", "1", "--segment-length", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == metric and rec["unit"] == "clips/s" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
