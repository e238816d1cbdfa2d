"""Turns the `ncu --page raw --csv` exports of tools/ncu_session_r02.sh into the small JSON summaries kept under profiles/r02/.
usage: python tools/ncu_summarise_r02.py gpurun_out profiles/r02"""
import csv, json, os, sys

src, dst = sys.argv[1], sys.argv[2]
KEYS = {
    "gpu__time_duration.sum": "gpu_time",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_hmma_pct",
    "sm__inst_executed_pipe_uniform.sum": "uniform_pipe_insts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed": "lts_sectors_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active",
}


def read_raw(path):
    """ncu raw page: one row per launch, one column per metric (two header lines: names, units)."""
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(l for l in fh if not l.startswith("=="))]
    if len(rows) < 3:
        return []
    names, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for n, u, v in zip(names, units, r):
            if n in ("Kernel Name",):
                d["kernel"] = v
            if n in KEYS:
                try:
                    d[KEYS[n]] = float(v.replace(",", ""))
                    d[KEYS[n] + "_unit"] = u
                except ValueError:
                    pass
        out.append(d)
    return out


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ms(v, unit):
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)


summary = {}
for name in sorted(os.listdir(src)):
    if not (name.startswith("r02_ncu_") and name.endswith("_raw.csv")):
        continue
    launches = read_raw(os.path.join(src, name))
    key = name[len("r02_ncu_"):-len("_raw.csv")]
    recs = []
    for d in launches:
        rec = {"kernel": d.get("kernel", "?")[:80]}
        if "gpu_time" in d:
            rec["gpu_time_ms"] = to_ms(d["gpu_time"], d["gpu_time_unit"])
        for k in ("dram_bytes_read", "dram_bytes_write"):
            if k in d:
                rec[k] = to_bytes(d[k], d[k + "_unit"])
        for k in ("dram_throughput_pct", "tensor_pipe_pct", "tensor_pipe_hmma_pct", "warps_active_pct", "issue_active_pct",
                  "lts_sectors_pct", "lts_throughput_pct", "l1tex_throughput_pct", "sm_throughput_pct", "registers_per_thread",
                  "grid_size", "block_size"):
            if k in d:
                rec[k] = d[k]
        recs.append(rec)
    summary[key] = recs
with open(os.path.join(dst, "ncu_summary_r02.json"), "w") as fh:
    json.dump(summary, fh, indent=1)
# decode megakernel traffic record (bench.py's `traffic`)
mega = summary.get("decode_mega", [])
if mega and "dram_bytes_read" in mega[0]:
    m = mega[0]
    steps, B, L0 = 16, 64, 514
    wbytes = 2.0 * (12 * (4 * 768 * 768 + 3 * 768 * 3072) + 16386 * 768)
    alg = steps * wbytes + sum(B * (L0 + i) * 12 * 2 * 768 * 2.0 for i in range(steps))
    rec = {"kernel": m["kernel"], "source": "profiles/r02/ncu_decode_mega_traffic.json",
           "capture": "ncu --set full --clock-control none, tools/mega_ncu.py NEW=17 (16 decode steps, B=64, prompt 514), round-2 build "
                      "(4 MMA-issuing warps, sequence-split attention for small batches)",
           "gpu_time_ms": m.get("gpu_time_ms"), "dram_bytes_read": m["dram_bytes_read"], "dram_bytes_write": m.get("dram_bytes_write", 0.0),
           "dram_bytes": m["dram_bytes_read"] + m.get("dram_bytes_write", 0.0), "decode_steps": steps, "algorithmic_bytes": alg,
           "ratio_to_algorithmic": (m["dram_bytes_read"] + m.get("dram_bytes_write", 0.0)) / alg,
           "issue_active_pct": m.get("issue_active_pct"), "dram_throughput_pct": m.get("dram_throughput_pct")}
    with open(os.path.join(dst, "ncu_decode_mega_traffic.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
print(json.dumps(summary, indent=1)[:3000])
