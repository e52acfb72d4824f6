#!/bin/bash
# e2e pipeline variants of bench.py (pack mode x streams x sub-batch x pack threads); short runs, one JSON line each.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
: > $out/${tag}_e2e_variants.jsonl
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --sustained-s 0 "$@" >> $out/${tag}_e2e_variants.jsonl 2>> $out/${tag}_e2e_variants.err; }
run --e2e-pack on --e2e-streams 2 --sub-batch 32
run --e2e-pack on --e2e-streams 3 --sub-batch 4
run --e2e-pack on --e2e-streams 4 --sub-batch 2
MRIACL_PACK_STREAM=0 run --e2e-pack on --e2e-streams 3 --sub-batch 4
MRIACL_PACK_STREAM=0 run --e2e-pack on --e2e-streams 4 --sub-batch 2
MRIACL_PACK_STREAM=0 run --e2e-pack on --e2e-streams 6 --sub-batch 1
MRIACL_PACK_STREAM=0 run --e2e-pack on --e2e-streams 2 --sub-batch 2
python - <<PY
import json
for l in open("$out/${tag}_e2e_variants.jsonl"):
    d = json.loads(l); e = d["e2e"]; m = e["modes"][e["mode"]]
    print(e["mode"], e["streams"], e["sub_batch"], round(e["value"]), "pack ms", round(m["host_pack_ms_last_step"], 1), "wait ms", round(m["host_wait_ms_last_step"], 1), "step ms", round(m["ms_per_step"], 1))
PY
