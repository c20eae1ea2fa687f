mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items()})"; }
export NNCONV_B200_Y_BYTES=150994944
echo "=== base"; run
echo "=== no red"; NNCONV_DEBUG=1 run
echo "=== stages 3"; NNCONV_CONV_STAGES=3 run
echo "=== stages 2"; NNCONV_CONV_STAGES=2 run
echo "=== stages 3 no red"; NNCONV_DEBUG=1 NNCONV_CONV_STAGES=3 run
