#!/bin/bash
# A/B of the experimental schedules (libmriacl_recon_exp.so): step time of configs[1] for each "ENV=... ENV=..." argument.
export MRIACL_RECON_LIBRARY=$PWD/mri_acl_imagesegmentation_adsp_b200/csrc/libmriacl_recon_exp.so
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-extra --sustained-s 0 --steps 30 --e2e-steps 1 --e2e-pack off 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().splitlines()[-1]); print('$cfg', round(d['ms_per_step'],4), d['e2e']['parity']['ok'])
except Exception as e: print('$cfg', 'FAILED', e)"
done
