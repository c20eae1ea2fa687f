mkdir -p gpurun_out
echo "=== bench241" ; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench241_r1a.json
echo "=== ncu launch list (darcy85, 1 step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1a.csv python bench.py --workload darcy85 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
echo "=== ncu full: conv + gemm kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 30 -c 2 -o gpurun_out/prof_conv_r1a -f python bench.py --workload darcy85 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 40 -c 3 -o gpurun_out/prof_gemm_r1a -f python bench.py --workload darcy85 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out
