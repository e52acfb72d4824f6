#!/usr/bin/env bash
# A/B helper for gpurun: one bench line per environment setting.
#   usage (inside gpurun):  bash tools/run_variants.sh "MRIACL_SCHEDULE=coresident" "MRIACL_RP16_CFG=2" ...
run() { env $1 timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_var.json 2> gpurun_out/b_var.err; python -c "
import json
d=json.load(open('gpurun_out/b_var.json'))
k=d['roofline']['kernels_ms']
print('$1', round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['frac'],4), {a:round(b,4) for a,b in k.items()})
" || tail -3 gpurun_out/b_var.err; }
if [ $# -eq 0 ]; then run "MRIACL_X=1"; fi
for v in "$@"; do run "$v"; done
