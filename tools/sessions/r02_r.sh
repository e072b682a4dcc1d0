#!/bin/bash
# Round 2, GPU session R: K* sparsity of the observation GP over the tiles the filter produces (measurement).
mkdir -p gpurun_out
timeout 600 python tools/kstar_sparsity.py > gpurun_out/kstar_sparsity_r02.json 2> gpurun_out/kstar_sparsity.err
tail -3 gpurun_out/kstar_sparsity.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/kstar_sparsity_r02.json'))
print(d['workload'])
for r in d['frames']:
    print(r['frame'], r['entries_below_1e-30'], {k:{t:[round(x,3) for x in v.values()] for t,v in r[k].items()} for k in ('filter_order','sorted','both_morton_sorted')})
PY
