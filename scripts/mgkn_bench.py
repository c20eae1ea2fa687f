#!/usr/bin/env python
"""SURVEY 8(f2) probe: MGKN V-cycle (neurips1_MGKN.py case 0: m=[2400,1600,400,100,25] on a 241^2 grid, width 64,
ker_width 256, depth 4) -- this library eager, this library replayed from a CUDA graph, and the reference-equivalent
torch path (oracle ops on CUDA tensors, fp32).  Prints ms per forward and edge-applications/s."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from graph_pde_b200 import graphs  # noqa: E402
from graph_pde_b200.models import KernelInduced  # noqa: E402
from oracle import nnconv_oracle as O  # noqa: E402  (baseline leg only)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    s, m = 241, [2400, 1600, 400, 100, 25]
    ri = [0.5 / 8 * 1.41, 0.5 / 8, 0.5 / 4, 0.5 / 2, 0.5]
    rx = [0.5 / 8 * 1.1, 0.5 / 8 * 1.41, 0.5 / 4 * 1.41, 0.5 / 2 * 1.41]
    theta = torch.randn(s * s)
    t0 = time.time()
    g = graphs.multi_level_ball_graph(s, m, ri, rx, theta=theta, device=dev)
    torch.cuda.synchronize()
    print('graph build on GPU: %.1f ms; edges mid/down/up = %d/%d/%d' %
          ((time.time() - t0) * 1e3, g.edge_index_mid.size(1), g.edge_index_down.size(1), g.edge_index_up.size(1)))
    depth, level = 4, len(m)
    model = KernelInduced(width=64, ker_width=256, depth=depth, ker_in=6, points=m, level=level, in_width=6).to(dev).eval()
    g.x = torch.randn(sum(m), 6, device=dev)
    edge_apps = depth * (g.edge_index_mid.size(1) + g.edge_index_down.size(1) + g.edge_index_up.size(1))
    with torch.no_grad():
        out = model(g)
        if os.environ.get('MGKN_PROFILE'):       # ncu --profile-from-start off: the launches of ONE eager forward
            model(g)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            model(g)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            return
        t_eager = timed(lambda: model(g), 20)
        print('this library, eager      : %8.3f ms  %.3e edge-apps/s' % (t_eager, edge_apps / t_eager * 1e3))
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    model(g)
            torch.cuda.current_stream().wait_stream(side)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                out_g = model(g)
            t_graph = timed(cg.replay, 50)
            err = float((out_g - out).abs().max() / out.abs().max())
            print('this library, CUDA graph : %8.3f ms  %.3e edge-apps/s  (replay vs eager rel diff %.1e)' %
                  (t_graph, edge_apps / t_graph * 1e3, err))
        except Exception as ex:  # noqa: BLE001
            print('CUDA graph capture failed:', repr(ex)[:300])
        # reference-equivalent torch path on the same device
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        p = {k: v.detach() for k, v in model.state_dict().items()}
        data = dict(edge_index_down=g.edge_index_down, edge_index_mid=g.edge_index_mid, edge_index_up=g.edge_index_up,
                    edge_attr_down=g.edge_attr_down, edge_attr_mid=g.edge_attr_mid, edge_attr_up=g.edge_attr_up,
                    range_down=g.edge_index_down_range.tolist(), range_mid=g.edge_index_range.tolist(),
                    range_up=g.edge_index_up_range.tolist())
        ref = O.mgkn_vcycle_forward(g.x, data, p, depth, level, m, variant='neurips1')
        print('parity vs reference-equivalent path: rel err %.2e' % float((out - ref).abs().max() / ref.abs().max()))
        t_ref = timed(lambda: O.mgkn_vcycle_forward(g.x, data, p, depth, level, m, variant='neurips1'), 5)
        print('reference-equivalent torch: %8.3f ms  %.3e edge-apps/s  -> speed-up eager %.1fx' %
              (t_ref, edge_apps / t_ref * 1e3, t_ref / t_eager))


if __name__ == '__main__':
    main()
