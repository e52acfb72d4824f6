#!/bin/bash
# A/B of the L2-persistence window on the tail of T (product library): step time per MRIACL_L2_PERSIST_MB
out=gpurun_out/${1}_persist.txt
: > $out
shift
for mb in "$@"; do
  MRIACL_L2_PERSIST_MB=$mb timeout 120 python tools/exp_time.py 2>&1 | tail -1 >> $out
done
cat $out
