run() { env "$@" timeout 100 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_var.json 2> gpurun_out/b_var.err; python -c "
import json,sys
d=json.load(open('gpurun_out/b_var.json'))
k=d['roofline']['kernels_ms']
print('$*', round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['frac'],4))
" || tail -3 gpurun_out/b_var.err; }
run MRIACL_RP_REVERSE=0
run MRIACL_RP_REVERSE=1
run MRIACL_RP_REVERSE=0
run MRIACL_RP_REVERSE=1
