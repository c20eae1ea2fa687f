mkdir -p gpurun_out
echo "=== tests" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for yb in 67108864 100663296 134217728; do
echo "=== bench241 Y_BYTES=$yb" ; NNCONV_B200_Y_BYTES=$yb timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step'], [ (r['kernel'], r['frac']) for r in d['roofline_kernels']])"
done
echo "=== bench85" ; timeout 600 python bench.py --workload darcy85 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], [ (r['kernel'], r['frac']) for r in d['roofline_kernels']])"
