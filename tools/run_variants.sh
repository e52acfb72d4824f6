run() { env "$@" timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_var.json 2> gpurun_out/b_var.err; python -c "
import json,sys
d=json.load(open('gpurun_out/b_var.json'))
k=d['roofline']['kernels_ms']
print('$*', 'col', round(k['colpass640_alone'],4))
" || tail -3 gpurun_out/b_var.err; }
run MRIACL_CP_DEBUG_SKIP=5
run MRIACL_CP_DEBUG_SKIP=3
run MRIACL_CP_DEBUG_SKIP=7
