"""Run under torch.distributed.run on N GPUs: BASELINE config 5's decomposition on the GPU -- the orthogonal MGKN
(MGKN_orthogonal_burgers1d.py) on the 1-D multipole hierarchy, s = 8192, 5 levels, cut into N aligned node ranges with
3-node halos per level (partition.Range1DPartition, one NCCL all-gather per depth iteration), CUDA convs of this
library -- against the unpartitioned forward of the same model on every rank.  Prints error and times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from graph_pde_b200 import graphs, partition
from graph_pde_b200.models import MGKN

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
s, levels, depth, width = 8192, 5, 4, 64
torch.manual_seed(0)
X, eis, eas = graphs.multi_pole_grid1d(torch.randn(s), s, is_periodic=True, levels=levels, device=dev)
model = MGKN(width=width, ker_width=1024, depth=depth, ker_in=4, in_width=2, s=s).to(dev).eval()
part = partition.Range1DPartition(s, levels, rank, world, halo=3, periodic=True)
ei_loc, ea_loc = [], []
for l, (ei, ea) in enumerate(zip(eis, eas)):
    e, m = part.local_edges(0 if l == 0 else l - 1, ei)
    ei_loc.append(e)
    ea_loc.append(ea[m].contiguous())
convs = [model.conv_list[l] for l in range(levels)] + [model.conv_list[-1]]      # set l -> conv l, last set -> conv -1
convs = [(lambda c: (lambda x, ei, ea: c(x, ei, ea)))(c) for c in convs]


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


with torch.no_grad():
    full = model((X, None, eis, eas))
    run = lambda: partition.partitioned_mgkn_forward(part, X[0][part.lo:part.hi], ei_loc, ea_loc, convs, model.fc1,  # noqa: E731
                                                     model.fc2, model.fc3, depth, width)
    out = run()
    err = float((out - full[part.lo:part.hi]).abs().max() / full.abs().max())
    ms_part = timed(run)
    ms_full = timed(lambda: model((X, None, eis, eas)))
print('rank %d/%d nodes [%d,%d) local edges %s: max rel err vs unpartitioned %.3e; partitioned %.3f ms, unpartitioned '
      '(one GPU) %.3f ms' % (rank, world, part.lo, part.hi, [int(e.size(1)) for e in ei_loc], err, ms_part, ms_full), flush=True)
assert err < 2e-3
dist.barrier()
dist.destroy_process_group()
