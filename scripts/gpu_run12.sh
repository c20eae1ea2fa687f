mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0})"; }
echo "=== ring2 nb512";  NNCONV_RING=2 NNCONV_B200_Y_BYTES=134217728 run
echo "=== ring2 nb1024"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=268435456 run
echo "=== ring2 nb2048"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=536870912 run
echo "=== ring2 nb8192"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=2147483648 run
echo "=== ncu fused (ring2 nb256)"
NNCONV_RING=2 NNCONV_B200_Y_BYTES=67108864 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_apply_tc -s 7 -c 1 -o gpurun_out/prof_apply_r1d -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_apply.log 2>&1
tail -2 gpurun_out/ncu_apply.log
