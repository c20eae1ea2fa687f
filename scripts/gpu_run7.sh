mkdir -p gpurun_out
export NNCONV_B200_Y_BYTES=100663296 NNCONV_NO_PIPE=1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 10 -c 2 -o gpurun_out/prof_conv_r1b -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv_b.log 2>&1
tail -2 gpurun_out/ncu_conv_b.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 3 -c 6 -o gpurun_out/prof_gemm_r1b -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm_b.log 2>&1
tail -2 gpurun_out/ncu_gemm_b.log
