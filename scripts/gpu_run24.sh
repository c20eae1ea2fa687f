mkdir -p gpurun_out
echo "=== backward tests" ; timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q 2>&1 | tail -25
echo "=== all gpu tests" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
