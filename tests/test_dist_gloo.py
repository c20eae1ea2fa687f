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


def _mgkn1d_setup():
    s, levels, width, depth = 128, 3, 8, 2
    torch.manual_seed(0)
    theta = torch.randn(s)
    X, eis, eas = graphs.multi_pole_grid1d(theta, s, is_periodic=True, levels=levels)
    p = {}
    gen = torch.Generator().manual_seed(1)
    p['fc1.weight'], p['fc1.bias'] = torch.randn(width, 2, generator=gen) * 0.5, torch.randn(width, generator=gen) * 0.1
    p['fc2.weight'], p['fc2.bias'] = torch.randn(16, width, generator=gen) * 0.3, torch.randn(16, generator=gen) * 0.1
    p['fc3.weight'], p['fc3.bias'] = torch.randn(1, 16, generator=gen) * 0.3, torch.randn(1, generator=gen) * 0.1
    for l in range(levels + 1):
        ws, bs, root, bias = O.reference_init(width, width, [4, 8, 8, width * width], seed=10 + l)
        for i, (w, b) in enumerate(zip(ws, bs)):
            p['conv_list.%d.nn.layers.%d.weight' % (l, 2 * i)] = w
            p['conv_list.%d.nn.layers.%d.bias' % (l, 2 * i)] = b
        p['conv_list.%d.root' % l], p['conv_list.%d.bias' % l] = root, bias
    return s, levels, width, depth, X, eis, eas, p


def _mgkn1d_job(rank, world):
    import torch.nn.functional as F
    s, levels, width, depth, X, eis, eas, p = _mgkn1d_setup()
    part = partition.Range1DPartition(s, levels, rank, world, halo=3, periodic=True)
    ei_loc, ea_loc = [], []
    for l, (ei, ea) in enumerate(zip(eis, eas)):
        j = 0 if l == 0 else l - 1
        e, m = part.local_edges(j, ei)
        ei_loc.append(e)
        ea_loc.append(ea[m])

    def mk(l):
        ws, bs = O.mlp_params_from_state(p, 'conv_list.%d.nn.' % l)
        return lambda x, ei, ea: O.nnconv_forward(x, ei, ea, ws, bs, p['conv_list.%d.root' % l], p['conv_list.%d.bias' % l], 'mean')
    convs = [mk(l) for l in range(levels + 1)]
    lin = lambda k: (lambda t: F.linear(t, p[k + '.weight'], p[k + '.bias']))   # noqa: E731
    out = partition.partitioned_mgkn_forward(part, X[0][part.lo:part.hi], ei_loc, ea_loc, convs, lin('fc1'), lin('fc2'),
                                             lin('fc3'), depth, width)
    return (part.lo, part.hi, out.numpy())


def test_range1d_partition_of_the_multipole_hierarchy_matches_unpartitioned():
    """BASELINE config 5's decomposition: aligned node ranges, 3-node halos per level, one all-gather per depth
    iteration -- partitioned MGKN forward == the oracle's unpartitioned forward."""
    s, levels, width, depth, X, eis, eas, p = _mgkn1d_setup()
    ref = O.mgkn_orthogonal_forward(X[0], eis, eas, p, depth, width, s).numpy()
    parts = _run(_mgkn1d_job, 2)
    got = np.zeros_like(ref)
    for lo, hi, out in parts:
        got[lo:hi] = out
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)
