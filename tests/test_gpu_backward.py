"""GPU: gradients of the CUDA backward (csrc/backward.cu through nnconv_backward) vs torch autograd through the
CPU oracle (SURVEY 4 "gradient parity": x, all MLP parameters, root, bias).  pytest -m gpu."""
import numpy as np
import pytest
import torch

from oracle import nnconv_oracle as O
from tests.helpers import DenseNetLike, make_conv

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _graph(gen, N, E, hub=False):
    src = torch.randint(0, N - 5, (E,), generator=gen)
    dst = torch.randint(2, N, (E,), generator=gen)
    if hub:
        src[:300] = 3
    order = torch.argsort(src, stable=True) if not hub else torch.arange(E)
    return torch.stack([src[order], dst[order]])


@pytest.mark.parametrize('layers,cin,cout,aggr,rw,bs,fwd_prec', [
    ([6, 32, 48, 16 * 16], 16, 16, 'mean', True, True, 'fp32'),
    ([6, 64, 64, 32 * 32], 32, 32, 'mean', True, True, 'f16'),      # tensor-core forward, fp32 backward
    ([4, 24, 5 * 7], 5, 7, 'add', False, True, 'fp32'),             # 2-layer MLP, odd shapes, no root
    ([3, 8 * 8], 8, 8, 'mean', True, False, 'fp32'),                # single Linear edge network
    ([6, 16, 32, 24, 64 * 64], 64, 64, 'mean', False, False, 'f16'),  # 4-layer MLP, MGKN style
])
def test_backward_matches_autograd_through_oracle(layers, cin, cout, aggr, rw, bs, fwd_prec, monkeypatch):
    from graph_pde_b200 import nn_conv
    from graph_pde_b200.nn_conv import NNConv_old
    monkeypatch.setattr(nn_conv, '_BWD_MODE', 'fp32')          # this file pins the CUDA-core fp32 backward
    gen = torch.Generator().manual_seed(17)
    N, E = 120, 1500
    ei = _graph(gen, N, E, hub=True)                 # unsorted sources + a hub with several tiles
    ea = torch.randn(E, layers[0], generator=gen)
    x = torch.randn(N, cin, generator=gen)
    torch.manual_seed(3)
    mlp = DenseNetLike(layers)
    lin = [m for m in mlp.layers if isinstance(m, torch.nn.Linear)]
    ws = [l.weight.detach().clone() for l in lin]
    bsl = [l.bias.detach().clone() for l in lin]
    root = torch.randn(cin, cout) * 0.2 if rw else None
    bias = torch.randn(cout) * 0.2 if bs else None
    gout = torch.randn(N, cout, generator=gen)
    # ---- reference gradients: autograd through the CPU oracle (fp64 for a clean reference)
    leaves = [t.double().requires_grad_(True) for t in [x] + ws + bsl + ([root] if rw else []) + ([bias] if bs else [])]
    xr, wr, br = leaves[0], leaves[1:1 + len(ws)], leaves[1 + len(ws):1 + 2 * len(ws)]
    rest = leaves[1 + 2 * len(ws):]
    rr = rest.pop(0) if rw else None
    bbr = rest.pop(0) if bs else None
    out_ref = O.nnconv_forward(xr, ei, ea.double(), wr, br, rr, bbr, aggr, cin, cout)
    (out_ref * gout.double()).sum().backward()
    # ---- CUDA
    conv = make_conv(NNConv_old, ws, bsl, root, bias, aggr, cin, cout, fwd_prec, DEV)
    xd = x.to(DEV).requires_grad_(True)
    out = conv(xd, ei.to(DEV), ea.to(DEV))
    (out * gout.to(DEV)).sum().backward()

    def chk(name, got, ref, tol=2e-4):
        ref = ref.float()
        err = float((got.detach().cpu() - ref).abs().max() / ref.abs().max().clamp(min=1e-20))
        assert err < tol, (name, err)

    chk('x', xd.grad, xr.grad)
    lin_d = [m for m in conv.nn.layers if isinstance(m, torch.nn.Linear)]
    for i, l in enumerate(lin_d):
        chk('W%d' % i, l.weight.grad, wr[i].grad)
        chk('b%d' % i, l.bias.grad, br[i].grad)
    if rw:
        chk('root', conv.root.grad, rr.grad)
    if bs:
        chk('bias', conv.bias.grad, bbr.grad)


def test_kernelnn_training_step_shares_gradients_over_T(monkeypatch):
    from graph_pde_b200 import nn_conv
    monkeypatch.setattr(nn_conv, '_BWD_MODE', 'fp32')
    """KernelNN applies ONE conv T times: parameter gradients must accumulate over the T applications
    (UAI1_full_resolution.py:29-30, loss.backward() :266)."""
    from graph_pde_b200.models import KernelNN
    gen = torch.Generator().manual_seed(5)
    s, r, w, kw, T = 9, 0.3, 16, 32, 3
    ei = torch.from_numpy(O.ball_connectivity(s, r))
    theta = np.random.RandomState(0).randn(s * s)
    ea = torch.from_numpy(O.ball_edge_attr(O.square_grid(s), ei.numpy(), theta))
    node_x = torch.randn(s * s, 6, generator=gen)
    y = torch.randn(s * s, 1, generator=gen)
    torch.manual_seed(0)
    model = KernelNN(w, kw, T, 6, in_width=6, precision='fp32').to(DEV)
    st = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
    # oracle + autograd (fp64)
    leaves = {k: v.clone().requires_grad_(True) for k, v in st.items()}
    out_ref = O.kernelnn_forward(node_x.double(), ei, ea.double(), leaves, T)
    torch.nn.functional.mse_loss(out_ref, y.double()).backward()

    class D(object):
        pass
    d = D()
    d.x, d.edge_index, d.edge_attr = node_x.to(DEV), ei.to(DEV), ea.to(DEV)
    loss = torch.nn.functional.mse_loss(model(d), y.to(DEV))
    loss.backward()
    for k, p in model.named_parameters():
        ref = leaves[k].grad.float()
        err = float((p.grad.cpu() - ref).abs().max() / ref.abs().max().clamp(min=1e-20))
        assert err < 5e-4, (k, err)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)      # the step of the reference loop (:271) works
    opt.step()
