#!/usr/bin/env python
"""One training step (forward + tensor-core backward + Adam) of KernelNN on a synthetic Darcy graph, bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly one step:

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/train_launches.csv python scripts/train_probe.py [darcy241|darcy85]
Without ncu it prints the CUDA-event time of the step."""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from graph_pde_b200 import graphs  # noqa: E402
from graph_pde_b200.models import KernelNN  # noqa: E402


def main():
    cfg = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'darcy241']
    dev = torch.device('cuda:0')
    s, r, w, kw, T = cfg['s'], cfg['r'], cfg['width'], cfg['ker_width'], cfg['depth']
    torch.manual_seed(0)
    model = KernelNN(w, kw, T, 6, in_width=6, precision='f16').to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    x6, ei, ea = graphs.darcy_sample(s, r, dev, seed=0)
    y = torch.randn(s * s, 1, device=dev)

    class D(object):
        pass
    d = D()
    d.x, d.edge_index, d.edge_attr = x6, ei, ea

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.norm(model(d).view(-1) - y.view(-1), 1)
        loss.backward()
        opt.step()
        return loss
    step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    a.record()
    step()
    b.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print('one training step: %.2f ms (E=%d, T=%d)' % (a.elapsed_time(b), ei.size(1), T))


if __name__ == '__main__':
    main()
