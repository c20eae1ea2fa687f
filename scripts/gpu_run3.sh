mkdir -p gpurun_out
echo "=== tests" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== bench85" ; timeout 600 python bench.py --workload darcy85 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], [ (r['kernel'], r['frac']) for r in d['roofline_kernels']])"
for yb in 50331648 100663296; do
echo "=== bench241 Y_BYTES=$yb" ; NNCONV_B200_Y_BYTES=$yb timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step'], [ (r['kernel'], r['frac']) for r in d['roofline_kernels']])"
done
echo "=== cpu baseline"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tail -1
