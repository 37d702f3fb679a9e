#!/bin/bash
# Process-to-process spread of the fused RMSD / Rg kernels: the `rmsd` and `rg` records of
# eight separate bench processes (256 headline frames per process keep the run short; the records
# themselves are the full configs[2] batches).  gpurun -- 'bash scripts/fused_spread.sh > gpurun_out/r2_fused_spread.json'
for k in 1 2 3 4 5 6 7 8; do
  python bench.py --frames 296 --steps 1 --warmup 3 --no-cpu-baseline --extras rmsd,rg 2>/dev/null
done | python -c "
import json, statistics, sys
runs = {'rmsd': [], 'rg': []}
clocks = {'rmsd': [], 'rg': []}
for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        continue
    for record in json.loads(line).get('extra', []):
        if record['name'] in runs:
            runs[record['name']].append(record['roofline']['frac'])
            clocks[record['name']].append(record['clocks'].get('sm_mhz'))
out = {}
for name, values in runs.items():
    out[name] = {'processes': len(values), 'frac_of_hbm_peak': {'min': min(values), 'median': statistics.median(values), 'max': max(values)},
                 'all': values, 'sm_mhz': clocks[name]}
print(json.dumps(out))
"
