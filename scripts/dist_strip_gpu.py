"""Run under torch.distributed.run on N GPUs: strip-partitioned conv stack (CUDA conv + NCCL halo all-gather)
vs the unpartitioned CUDA result computed on every rank.  Prints max relative error per rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from graph_pde_b200 import graphs, partition
from graph_pde_b200.models import KernelNN

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
s, r, w, kw, T = 85, 0.10, 64, 256, 4
torch.manual_seed(0)
model = KernelNN(w, kw, T, 6, in_width=6).to(dev).eval()
x6, ei, ea = graphs.darcy_sample(s, r, dev, seed=0)
with torch.no_grad():
    x0 = model.fc1(x6)
    full = model.conv_stack(x0, ei, ea)
    part = partition.StripPartition(s, r, rank, world, device=dev)
    grid = graphs.square_grid(s, dev)
    ea_loc = graphs.ball_edge_attr(grid, part.edge_index_global, x6[:, 2])
    conv = lambda xl, e, a: model.conv1(xl, e, a)   # noqa: E731
    out = partition.partitioned_conv_stack(conv, part.local_slice(x0).clone(), part, ea_loc, T)
    halo = partition.PeerHalo(part, w, dev)
    outs_p = [partition.partitioned_conv_stack_peer(conv, part.local_slice(x0).clone(), part, ea_loc, T, halo)
              for _ in range(3)]                        # repeated: sequence flags keep counting across stacks
ref = full[part.row_lo * s:part.row_hi * s]
err = float((out - ref).abs().max() / ref.abs().max())
err_p = max(float((o - ref).abs().max() / ref.abs().max()) for o in outs_p)
print('rank %d/%d rows [%d,%d) local edges %d  max rel err vs unpartitioned: all-gather %.3e, peer push %.3e' %
      (rank, world, part.row_lo, part.row_hi, part.edge_index.size(1), err, err_p), flush=True)
assert err < 2e-3 and err_p < 2e-3
halo.close()
dist.barrier()
dist.destroy_process_group()
