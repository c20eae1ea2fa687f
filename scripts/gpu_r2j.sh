#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward_tc.py tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -q > $O/r2j_tests.log 2>&1; tail -3 $O/r2j_tests.log
python - <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import bench
from graph_pde_b200 import nn_conv
dev = torch.device('cuda:0')
for mode, thr in (('auto', 16384), ('auto', 4096), ('off', 0)):
    nn_conv._EDGE_KERNELS = mode
    nn_conv._EDGE_KERNELS_MAX_EDGES = thr
    c = bench.other_configs(dev, 'f16')
    print('MODE', mode, thr, {k: {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ('ms_per_step', 'ms_per_forward', 'ms_per_forward_cuda_graph', 'nnconv_launches_per_forward')} for k, v in c.items()}, {k: v['parity']['max_rel_err'] for k, v in c.items()})
PY
python scripts/train_probe.py darcy241 | tail -1
