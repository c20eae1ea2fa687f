#!/usr/bin/env python
"""BASELINE config 5 probe: orthogonal MGKN on the 1-D multipole hierarchy (MGKN_orthogonal_burgers1d.py), s=8192,
5 levels, width 64, ker_width 1024, depth 4 -- this library eager / CUDA-graph replay vs the reference-equivalent
torch path (oracle ops on CUDA tensors, fp32)."""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from graph_pde_b200 import GraphedForward, graphs  # noqa: E402
from graph_pde_b200.models import MGKN  # noqa: E402
from oracle import nnconv_oracle as O  # noqa: E402  (baseline leg only)
from scripts.mgkn_bench import timed  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    s, levels, depth, width = 8192, 5, 4, 64
    X, ei, ea = graphs.multi_pole_grid1d(torch.randn(s), s, is_periodic=True, levels=levels, device=dev)
    model = MGKN(width=width, ker_width=1024, depth=depth, ker_in=4, in_width=2, s=s).to(dev).eval()
    data = (X, None, ei, ea)
    edge_apps = depth * sum(e.size(1) for e in ei)
    with torch.no_grad():
        out = model(data)
        t_eager = timed(lambda: model(data), 20)
        print('edge sets %s -> %d edge-apps per forward' % ([e.size(1) for e in ei], edge_apps))
        print('this library, eager      : %8.3f ms  %.3e edge-apps/s' % (t_eager, edge_apps / t_eager * 1e3))
        gf = GraphedForward(model, data)
        t_graph = timed(gf.replay, 50)
        print('this library, CUDA graph : %8.3f ms  %.3e edge-apps/s' % (t_graph, edge_apps / t_graph * 1e3))
        torch.backends.cuda.matmul.allow_tf32 = False
        p = {k: v.detach() for k, v in model.state_dict().items()}
        ref = O.mgkn_orthogonal_forward(X[0], ei, ea, p, depth, width, s)
        print('parity vs reference-equivalent path: rel err %.2e' % float((out - ref).abs().max() / ref.abs().max()))
        t_ref = timed(lambda: O.mgkn_orthogonal_forward(X[0], ei, ea, p, depth, width, s), 5)
        print('reference-equivalent torch: %8.3f ms  %.3e edge-apps/s  -> speed-up eager %.1fx, replay %.1fx' %
              (t_ref, edge_apps / t_ref * 1e3, t_ref / t_eager, t_ref / t_graph))


if __name__ == '__main__':
    main()
