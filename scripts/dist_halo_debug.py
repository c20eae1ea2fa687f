"""Debug aid (torchrun, N GPUs, CUDA_LAUNCH_BLOCKING=1): exercises the peer-store halo exchange step by step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from graph_pde_b200 import partition

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)


def say(*a):
    print('[rank %d]' % rank, *a, flush=True)


s, r, C = 85, 0.10, 64
part = partition.StripPartition(s, r, rank, world, device=dev)
say('part rows', part.row_lo, part.row_hi, 'n_local', part.n_local, 'own', part.own_lo, part.own_hi, 'ranges',
    partition.halo_ranges(part))
torch.manual_seed(0)
xg = torch.randn(s * s, C, device=dev)
x_loc = part.local_slice(xg).clone()
torch.cuda.synchronize()
halo = partition.PeerHalo(part, C, dev)
torch.cuda.synchronize()
say('PeerHalo built; base', hex(halo.base), 'up', halo.up, 'down', halo.down)
x = halo.load(x_loc)
torch.cuda.synchronize()
say('load ok')
out = x * 2.0 - 1.0                       # stands in for one application: owned rows are what matters
nxt = halo.advance(out, 0, relu=True)
torch.cuda.synchronize()
say('advance ok')
dist.barrier()
ref = torch.relu(part.local_slice(xg) * 2.0 - 1.0)
err = float((nxt - ref).abs().max())
say('halo rows after one advance: max abs err vs global =', err)
halo.finish()
torch.cuda.synchronize()
say('finish ok')
halo.close()
dist.barrier()
dist.destroy_process_group()
