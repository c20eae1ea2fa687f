mkdir -p gpurun_out
NNCONV_BENCH_WORKLOAD=darcy85 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_mlp12_tc -s 3 -c 1 -o gpurun_out/prof_mlp12_r1e -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp12.log 2>&1
tail -2 gpurun_out/ncu_mlp12.log
