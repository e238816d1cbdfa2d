"""Compact view of gpurun_out/summary.txt (bench lines shortened)."""
import json, sys
for line in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/summary.txt").read().splitlines():
    if line.startswith('{'):
        try:
            d = json.loads(line)
        except Exception:
            print("unparsed", line[:300]); continue
        if 'metric' in d:
            r = d.get('roofline', {})
            print(d['config']['workload'][:28], '| value', round(d['value'], 1), d['unit'], '| ms', round(d['ms_per_step'], 1), '|',
                  {k: round(v, 1) for k, v in d.get('stages_ms', {}).items()}, '| roof', r.get('kernel', '')[:18], round(r.get('frac', 0), 3),
                  {k: (round(v['frac'], 3), round(v['kernel_ms_per_step'], 1)) for k, v in r.get('other', {}).items()},
                  '| e2e', round((d.get('e2e') or {}).get('value') or 0, 1))
        else:
            for k, v in d.items():
                print(' ', k, json.dumps(v))
    else:
        print(line[:300])
