#!/bin/bash
# A/B of the TMA column pass / TMA co-resident kernel (experimental library): parity tests under "$1" env, then step times.
export MRIACL_RECON_LIBRARY=$PWD/mri_acl_imagesegmentation_adsp_b200/csrc/libmriacl_recon_exp.so
out=gpurun_out/${1}_probe.txt
: > $out
env $2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 >> $out
shift; shift
for cfg in "$@"; do
  env $cfg timeout 120 python tools/exp_time.py 2>&1 | tail -2 >> $out
done
cat $out
