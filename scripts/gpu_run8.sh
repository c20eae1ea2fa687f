mkdir -p gpurun_out
echo "=== tests" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items()})"; }
for yb in 100663296 150994944 201326592; do
echo "=== bench241 pipelined Y_BYTES=$yb"; NNCONV_B200_Y_BYTES=$yb run
done
echo "=== no pipe 96MB"; NNCONV_NO_PIPE=1 NNCONV_B200_Y_BYTES=100663296 run
echo "=== one per sm, pipelined 144MB"; NNCONV_CONV_ONE_PER_SM=1 NNCONV_B200_Y_BYTES=150994944 run
echo "=== bench85"; NNCONV_BENCH_WORKLOAD=darcy85 run
