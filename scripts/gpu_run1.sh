mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "=== kernels" ; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -25
echo "=== parity fp32" ; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fp32 or library" 2>&1 | tail -25
echo "=== parity rest" ; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not fp32" 2>&1 | tail -40
echo "=== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== bench85" ; timeout 600 python bench.py --workload darcy85 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -5
