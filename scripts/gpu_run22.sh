mkdir -p gpurun_out
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks']['sm_mhz'], [(r['kernel'][:10], round(r['frac'],3)) for r in d['roofline_kernels']])"; }
echo "=== default (ring2 nb256)"; run
echo "=== ring2 nb128"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=33554432 run
echo "=== ring3 nb128"; NNCONV_RING=3 NNCONV_B200_Y_BYTES=50331648 run
echo "=== ring4 nb64"; NNCONV_RING=4 NNCONV_B200_Y_BYTES=33554432 run
echo "=== darcy85"; NNCONV_BENCH_WORKLOAD=darcy85 run
