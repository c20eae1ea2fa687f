"""CPU, world_size = 2 over gloo: the host logic of the two multi-GPU decompositions run with
the CPU oracle standing in for the CUDA conv -- partitioned result == unpartitioned result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graph_pde_b200 import graphs, partition
from oracle import nnconv_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


S, R_, W, T = 21, 0.13, 8, 3


def _setup():
    torch.manual_seed(0)
    ws, bs, root, bias = O.reference_init(W, W, [6, 16, 16, W * W], seed=0)
    grid = graphs.square_grid(S)
    theta = torch.randn(S * S)
    x = torch.randn(S * S, W)
    return ws, bs, root, bias, grid, theta, x


def _strip_job(rank, world):
    ws, bs, root, bias, grid, theta, x = _setup()
    part = partition.StripPartition(S, R_, rank, world)
    ea = graphs.ball_edge_attr(grid, part.edge_index_global, theta)
    conv = lambda xl, ei, ea_: O.nnconv_forward(xl, ei, ea_, ws, bs, root, bias, 'mean')   # noqa: E731
    out = partition.partitioned_conv_stack(conv, part.local_slice(x).clone(), part, ea, T)
    return (part.row_lo * S, part.row_hi * S, out.numpy())


def test_strip_partition_with_halo_exchange_matches_unpartitioned():
    ws, bs, root, bias, grid, theta, x = _setup()
    ei = graphs.ball_connectivity(S, R_)
    ea = graphs.ball_edge_attr(grid, ei, theta)
    ref = O.kernelnn_conv_stack(x, ei, ea, ws, bs, root, bias, T).numpy()
    parts = _run(_strip_job, 2)
    got = np.zeros_like(ref)
    for lo, hi, out in parts:
        got[lo:hi] = out
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_strip_partition_edges_cover_the_graph_once():
    ei = graphs.ball_connectivity(S, R_)
    seen = []
    for rank in range(3):
        part = partition.StripPartition(S, R_, rank, 3)
        assert int(part.edge_index.min()) >= 0 and int(part.edge_index.max()) < part.n_local
        seen.append(part.edge_index_global)
    allp = torch.cat(seen, dim=1)
    assert allp.size(1) == ei.size(1)
    key = lambda e: (e[0] * (S * S) + e[1]).sort().values    # noqa: E731
    assert torch.equal(key(allp), key(ei))


def _batch_job(rank, world):
    ws, bs, root, bias, grid, theta, x = _setup()
    ei = graphs.ball_connectivity(S, R_)
    outs = {}
    lin = torch.nn.Linear(W, 1)
    torch.manual_seed(1)
    lin.reset_parameters()
    for gidx in partition.shard_indices(5, rank, world):
        th = torch.randn(S * S, generator=torch.Generator().manual_seed(100 + gidx))
        ea = graphs.ball_edge_attr(grid, ei, th)
        o = O.kernelnn_conv_stack(x, ei, ea, ws, bs, root, bias, 2)
        outs[gidx] = float(o.sum())
        lin(o).sum().backward()                      # accumulate local gradients
    partition.allreduce_gradients(lin)
    return (outs, lin.weight.grad.numpy().copy())


def test_batch_sharding_and_gradient_allreduce():
    res = _run(_batch_job, 2)
    merged = {}
    for outs, _ in res:
        merged.update(outs)
    assert sorted(merged) == [0, 1, 2, 3, 4]
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-6)          # same summed gradient on both ranks
    # serial reference of the gradient
    ws, bs, root, bias, grid, theta, x = _setup()
    ei = graphs.ball_connectivity(S, R_)
    lin = torch.nn.Linear(W, 1)
    torch.manual_seed(1)
    lin.reset_parameters()
    for gidx in range(5):
        th = torch.randn(S * S, generator=torch.Generator().manual_seed(100 + gidx))
        ea = graphs.ball_edge_attr(grid, ei, th)
        lin(O.kernelnn_conv_stack(x, ei, ea, ws, bs, root, bias, 2)).sum().backward()
    np.testing.assert_allclose(res[0][1], lin.weight.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_peer_halo_row_ranges_reproduce_the_all_gather_exchange():
    """halo_ranges (what PeerHalo pushes into the neighbours' buffers on the GPU) moves exactly the rows
    halo_exchange moves: simulate the pushes with plain copies between per-rank tensors."""
    world, C = 4, 3
    torch.manual_seed(1)
    xg = torch.randn(S * S, C)
    parts = [partition.StripPartition(S, R_, r, world) for r in range(world)]
    local = [torch.zeros(p.n_local, C) for p in parts]
    for p, x in zip(parts, local):
        x[p.own_lo:p.own_hi] = xg[p.row_lo * S:p.row_hi * S]          # only owned rows are known after an application
    for p, x in zip(parts, local):
        rg = partition.halo_ranges(p)
        if rg['up'] is not None:
            src, dst, n = rg['up']
            local[p.rank - 1][dst:dst + n] = x[src:src + n]
        if rg['down'] is not None:
            src, dst, n = rg['down']
            local[p.rank + 1][dst:dst + n] = x[src:src + n]
    for p, x in zip(parts, local):
        assert torch.equal(x, p.local_slice(xg)), p.rank
