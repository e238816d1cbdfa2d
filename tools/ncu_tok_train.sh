#!/bin/bash
# ncu launch list of ONE tokenizer training step (bench.py --workload train-tokenizer64 --batch 4): per-kernel serialised time.
# Never a bench value.  Output: gpurun_out/launches_tok_train.csv (summarise with tools/summarise_launches.py).
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SKIP=${SKIP:-7200}      # 3 warm-up steps x ~2400 launches
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP --launch-count 2600 --csv \
    --log-file gpurun_out/launches_tok_train.csv python bench.py --workload train-tokenizer64 --batch 4 --steps 1 --warmup 3 \
    > gpurun_out/launches_tok_train.out 2>&1 || true
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_tok_train.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
kn, mv = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
unit = rows[hdr].index("Metric Unit")
t = collections.Counter(); n = collections.Counter()
for r in rows[hdr + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[unit] in ("ns", "nsecond") else v     # -> us
    name = r[kn].split("(")[0].split("<")[0]
    t[name] += v; n[name] += 1
tot = sum(t.values())
with open("gpurun_out/launches_tok_train.txt", "w") as fh:
    fh.write(f"tokenizer training step (4 clips of 64x64x16), serialised kernel time {tot/1e3:.1f} ms over {sum(n.values())} launches\n")
    for k, v in t.most_common(25):
        fh.write(f"{v/1e3:9.2f} ms {100*v/tot:5.1f} %  x{n[k]:5d}  {k}\n")
print(open("gpurun_out/launches_tok_train.txt").read())
PY
