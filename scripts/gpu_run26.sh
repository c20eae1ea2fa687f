mkdir -p gpurun_out
for n in 8 4; do
echo "=== bench N=$n"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e6,1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,1), d['clocks'])"
done
echo "=== reference arm under torchrun N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-160
