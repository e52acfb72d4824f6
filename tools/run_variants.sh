for sb in 4 8 16 32; do timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 5 --sub-batch $sb > gpurun_out/b_var.json 2> gpurun_out/b_var.err; python -c "
import json
d=json.load(open('gpurun_out/b_var.json'))
e=d['e2e']; print('sub_batch', $sb, round(e['value'],1), 'slices/s', round(e['ms_per_step'],2), 'ms', round(e['h2d_bytes_per_step']/e['ms_per_step']/1e6,1), 'GB/s')
"; done
