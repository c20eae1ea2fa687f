mkdir -p gpurun_out
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items()}, [(r['kernel'][:10], round(r['frac'],3)) for r in d['roofline_kernels']])"; }
for yb in 67108864 134217728; do
echo "=== bench241 fused Y_BYTES=$yb ring4"; NNCONV_B200_Y_BYTES=$yb run
done
echo "=== bench241 fused 96MB ring3"; NNCONV_RING=3 NNCONV_B200_Y_BYTES=100663296 run
echo "=== bench241 fused 64MB ring2"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=67108864 run
echo "=== bench85 fused 64MB"; NNCONV_B200_Y_BYTES=67108864 NNCONV_BENCH_WORKLOAD=darcy85 run
