mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks']['sm_mhz'])"; }
echo "=== ring2 nb128 (32MB)"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=33554432 run
echo "=== ring2 nb256 (64MB)"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=67108864 run
echo "=== ring2 nb384 (96MB)"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=100663296 run
echo "=== ring3 nb128 (48MB)"; NNCONV_RING=3 NNCONV_B200_Y_BYTES=50331648 run
echo "=== darcy85 fused12"; NNCONV_BENCH_WORKLOAD=darcy85 run
echo "=== darcy85 unfused12"; NNCONV_NO_FUSE12=1 NNCONV_BENCH_WORKLOAD=darcy85 run
echo "=== darcy16"; NNCONV_BENCH_WORKLOAD=darcy16 run
