#!/bin/bash
# One GPU-box visit: parity tests, the contract bench, the ncu launch list and one full capture of a 64-slice step.
# usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_gputest.log
tail -3 $out/${tag}_gputest.log
timeout 600 python bench.py --steps 50 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
tail -c 600 $out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_raw.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:colpass|rowpass|normalize' -s 3 -c 3 -f -o $out/${tag}_step \
  python tools/profile_step.py --batch 64 --steps 2 --chunk 64 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $out
