mkdir -p gpurun_out
echo "=== tests" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench default"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench241_r1f.json; python -c "
import json; d=json.load(open('gpurun_out/bench241_r1f.json')); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],2), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks'], [(r['kernel'][:10], round(r['frac'],3)) for r in d['roofline_kernels']], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['gpu_launches'])"
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:90])"
echo "=== ncu launch list (darcy85)"
NNCONV_BENCH_WORKLOAD=darcy85 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_f.log 2>&1
tail -1 gpurun_out/ncu_launch_f.log | cut -c1-200
echo "=== ncu full apply"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_apply_tc -s 7 -c 1 -o gpurun_out/prof_apply_r1f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_apply_f.log 2>&1
tail -1 gpurun_out/ncu_apply_f.log | cut -c1-120
echo "=== ncu full gemm"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 4 -c 4 -o gpurun_out/prof_gemm_r1f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm_f.log 2>&1
tail -1 gpurun_out/ncu_gemm_f.log | cut -c1-120
