mkdir -p gpurun_out
echo "=== strip partition on 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_strip_gpu.py 2>&1 | grep -E "rank|Error|error" | head
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0})"
echo "=== bench N=1 (same box)"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks'], [(r['kernel'][:10], round(r['frac'],3)) for r in d['roofline_kernels']])"
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
