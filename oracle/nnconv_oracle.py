"""CPU oracle for the graph-pde NNConv hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (torch-CPU / numpy, fp32 or fp64) of the reference algorithm, each function citing
the reference file:line it follows (paths relative to /root/reference, GNO = graph-neural-operator,
MGNO = multipole-graph-neural-operator).  Only tests/, __graft_entry__.smoke() and bench.py's baseline
legs (cpu_baseline, --impl reference, and gpu_reference_port = these same torch ops on CUDA tensors, the
"reference single-GPU path" of SURVEY 8(d)(ii)) may import this module; the product path (graph_pde_b200)
never does and fails loudly when its CUDA library is missing.

Pinning: this restatement is checked (tests/test_oracle_golden.py) against golden vectors produced by
running the reference's own nn_conv.py / utilities.py UNMODIFIED in the build container
(oracle/gen_golden.py -> tests/golden/*.npz).  The third-party half of the path
(torch_geometric.MessagePassing.propagate + torch_scatter, not vendored, no version pinned by the
reference) is itself a restatement (oracle/pyg_stub), so parity is pinned against the reference's
files but "unpinned" against PyG/torch_scatter; the two judgement calls are mean(empty)=0 and an
unspecified summation order.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Edge MLP (GNO/utilities.py:201-227  DenseNet: Linear -> ReLU -> ... -> Linear, no out nonlinearity)
# ----------------------------------------------------------------------------------------------
def dense_net(edge_attr, weights, biases):
    """weights[l]: [out_l, in_l] (torch.nn.Linear layout), biases[l]: [out_l].
    GNO/utilities.py:223-227 applies the layers in order; ReLU after every layer but the last
    (ctor :212-221, nonlinearity=torch.nn.ReLU at every call site)."""
    h = edge_attr
    n = len(weights)
    for l in range(n):
        h = F.linear(h, weights[l], biases[l])
        if l != n - 1:
            h = torch.relu(h)
    return h


# ----------------------------------------------------------------------------------------------
# NNConv_old.forward / message / update  (GNO/nn_conv.py:267-282) + PyG propagate (not vendored)
# ----------------------------------------------------------------------------------------------
def nnconv_forward(x, edge_index, edge_attr, weights, biases, root=None, bias=None, aggr='mean',
                   in_channels=None, out_channels=None, edge_chunk=None):
    """out[n] = reduce_{e: dst_e = n} ( x[src_e] @ K_e ) + x[n] @ root + bias,
    K_e = DenseNet(edge_attr_e).view(in, out)   (GNO/nn_conv.py:273-275, row-major [in,out], H5)
    flow = source_to_target: src = edge_index[0], dst = edge_index[1]  (H1)
    reduce = add | mean (sum / max(deg,1), H4: mean of the empty set is 0).
    edge_chunk: process edges in chunks (exact up to fp summation order) so that [E, in*out] is never
    materialised for big graphs; None = one shot exactly like the reference."""
    x = x.unsqueeze(-1) if x.dim() == 1 else x                       # nn_conv.py:269
    pseudo = edge_attr.unsqueeze(-1) if edge_attr.dim() == 1 else edge_attr   # nn_conv.py:270
    n = x.size(0)
    cin = x.size(1) if in_channels is None else in_channels
    cout = (weights[-1].size(0) // cin) if out_channels is None else out_channels
    src, dst = edge_index[0], edge_index[1]
    e_total = src.numel()
    out = torch.zeros(n, cout, dtype=x.dtype, device=x.device)
    step = e_total if (edge_chunk is None or edge_chunk <= 0) else edge_chunk
    for s in range(0, max(e_total, 1), max(step, 1)):
        sl = slice(s, min(s + step, e_total))
        if sl.stop <= sl.start:
            break
        x_j = x.index_select(0, src[sl])                               # PyG propagate gather
        w = dense_net(pseudo[sl], weights, biases).view(-1, cin, cout)  # nn_conv.py:274
        msg = torch.matmul(x_j.unsqueeze(1), w).squeeze(1)              # nn_conv.py:275
        out.index_add_(0, dst[sl], msg)                                # PyG scatter_('add'|'mean')
    if aggr == 'mean':
        cnt = torch.zeros(n, dtype=x.dtype, device=x.device)
        cnt.index_add_(0, dst, torch.ones(e_total, dtype=x.dtype, device=x.device))
        out = out / cnt.clamp(min=1).unsqueeze(-1)
    elif aggr != 'add':
        raise ValueError('oracle supports aggr in {add, mean}')
    if root is not None:                                               # nn_conv.py:278-279
        out = out + torch.mm(x, root)
    if bias is not None:                                               # nn_conv.py:280-281
        out = out + bias
    return out


# ----------------------------------------------------------------------------------------------
# Callers
# ----------------------------------------------------------------------------------------------
def kernelnn_conv_stack(x, edge_index, edge_attr, weights, biases, root, bias, depth, aggr='mean',
                        relu_last=True, edge_chunk=None):
    """The T-loop of KernelNN.forward (GNO/UAI1_full_resolution.py:29-30): the SAME conv applied
    `depth` times with ReLU after each; relu_last=False reproduces GNO/UAI7_evaluate.py:29-32."""
    for k in range(depth):
        x = nnconv_forward(x, edge_index, edge_attr, weights, biases, root, bias, aggr,
                           edge_chunk=edge_chunk)
        if relu_last or k != depth - 1:
            x = torch.relu(x)
    return x


def kernelnn_forward(node_x, edge_index, edge_attr, p, depth, relu_last=True, edge_chunk=None):
    """KernelNN.forward (GNO/UAI1_full_resolution.py:26-33): fc1, T x relu(conv1), fc2.
    p: dict with fc1.weight, fc1.bias, conv1.nn.layers.{0,2,4}.{weight,bias}, conv1.root, conv1.bias,
    fc2.weight, fc2.bias (state-dict names of the reference module)."""
    ws, bs = mlp_params_from_state(p, 'conv1.nn.')
    x = F.linear(node_x, p['fc1.weight'], p['fc1.bias'])
    x = kernelnn_conv_stack(x, edge_index, edge_attr, ws, bs, p.get('conv1.root'), p.get('conv1.bias'),
                            depth, 'mean', relu_last, edge_chunk)
    return F.linear(x, p['fc2.weight'], p['fc2.bias'])


def mlp_params_from_state(state, prefix):
    """DenseNet state-dict keys are '<prefix>layers.<2l>.weight' (ReLU modules take the odd slots,
    GNO/utilities.py:212-218)."""
    idx = sorted({int(k[len(prefix) + 7:].split('.')[0]) for k in state
                  if k.startswith(prefix + 'layers.') and k.endswith('.weight')})
    return ([state['%slayers.%d.weight' % (prefix, i)] for i in idx],
            [state['%slayers.%d.bias' % (prefix, i)] for i in idx])


def mgkn_vcycle_forward(x_in, data, p, depth, level, points, variant='neurips1'):
    """KernelInduced.forward (MGNO/neurips1_MGKN.py:65-89) or MKGN.forward
    (MGNO/MGKN_general_darcy2d.py:69-94, variant='general').
    data: dict with edge_index_{down,mid,up}, edge_attr_{down,mid,up}, range_{down,mid,up} ([L,2]).
    p: state dict (fc_in, conv_down_list.l, conv_list.l, conv_up_list.l, fc_out1, fc_out2)."""
    def conv(prefix, x, ei, ea):
        ws, bs = mlp_params_from_state(p, prefix + 'nn.')
        return nnconv_forward(x, ei, ea, ws, bs, p.get(prefix + 'root'), p.get(prefix + 'bias'), 'mean')

    x = F.linear(x_in, p['fc_in.weight'], p['fc_in.bias'])
    rd, rm, ru = data['range_down'], data['range_mid'], data['range_up']
    pts = [0] + list(np.cumsum(points)) if variant == 'general' else None
    for _ in range(depth):
        for l in range(level - 1):                                     # :74-76 downward
            a, b = int(rd[l][0]), int(rd[l][1])
            x = torch.relu(x + conv('conv_down_list.%d.' % l, x, data['edge_index_down'][:, a:b],
                                    data['edge_attr_down'][a:b]))
        for l in reversed(range(level)):                               # :79-84 upward
            a, b = int(rm[l][0]), int(rm[l][1])
            if variant == 'general':                                   # MGKN_general_darcy2d.py:83-86
                lo, hi = int(pts[l]), int(pts[l + 1])
                x = x.clone()
                x[lo:hi] = conv('conv_list.%d.' % l, x[lo:hi].clone(),
                                data['edge_index_mid'][:, a:b] - lo, data['edge_attr_mid'][a:b])
            else:
                x = torch.relu(x + conv('conv_list.%d.' % l, x, data['edge_index_mid'][:, a:b],
                                        data['edge_attr_mid'][a:b]))
            if l > 0:
                a, b = int(ru[l - 1][0]), int(ru[l - 1][1])
                x = torch.relu(x + conv('conv_up_list.%d.' % (l - 1), x, data['edge_index_up'][:, a:b],
                                        data['edge_attr_up'][a:b]))
    n_out = points[0] if variant != 'general' else int(pts[1])
    x = torch.relu(F.linear(x[:n_out], p['fc_out1.weight'], p['fc_out1.bias']))
    return F.linear(x, p['fc_out2.weight'], p['fc_out2.bias'])


def mgkn_orthogonal_forward(X0, edge_index_list, edge_attr_list, p, depth, width, s):
    """MGKN.forward (MGNO/MGKN_orthogonal_burgers1d.py:59-86): avg_pool1d restriction, nearest
    upsample prolongation, NNConv (root + bias) per level."""
    def conv(l, x, ei, ea):
        prefix = 'conv_list.%d.' % l
        ws, bs = mlp_params_from_state(p, prefix + 'nn.')
        return nnconv_forward(x, ei, ea, ws, bs, p.get(prefix + 'root'), p.get(prefix + 'bias'), 'mean')

    def down(x, s_l):                                                  # :52-56
        x = x.transpose(0, 1).reshape(1, width, s_l)
        x = F.avg_pool1d(x, kernel_size=2)
        return x.reshape(width, -1).transpose(0, 1)

    def up(x, s_l):                                                    # :45-49
        x = x.transpose(0, 1).reshape(1, width, s_l)
        x = F.interpolate(x, scale_factor=2, mode='nearest')
        return x.reshape(width, -1).transpose(0, 1)

    level = len(edge_index_list) - 1
    nconv = len({k.split('.')[1] for k in p if k.startswith('conv_list.')})
    x = F.linear(X0, p['fc1.weight'], p['fc1.bias'])
    phi = [None] * level
    for _ in range(depth):
        for l in range(level):                                         # :67-71
            phi[l] = x
            if l != level - 1:
                x = down(x, s // (2 ** l))
        x = torch.relu(x + conv(nconv - 1, phi[-1], edge_index_list[-1], edge_attr_list[-1]))   # :74
        for l in reversed(range(level)):
            if l != 0:
                x = up(x, s // (2 ** l))                               # :78
                x = torch.relu(x + conv(l, phi[l - 1], edge_index_list[l], edge_attr_list[l]))   # :80
            else:
                x = torch.relu(x + conv(0, phi[0], edge_index_list[0], edge_attr_list[0]))       # :82
    x = torch.relu(F.linear(x, p['fc2.weight'], p['fc2.bias']))
    return F.linear(x, p['fc3.weight'], p['fc3.bias'])


# ----------------------------------------------------------------------------------------------
# Input side of the path: ball graph of a square mesh (GNO/utilities.py:229-285), restated WITHOUT
# the dense O(N^2) distance matrix.  Edge order = np.where(pwd <= r) order = src-major, dst ascending.
# ----------------------------------------------------------------------------------------------
def square_grid(s):
    """GNO/utilities.py:241-248: np.meshgrid default ('xy') indexing, x fastest: node = iy*s + ix,
    pos = (ix/(s-1), iy/(s-1))."""
    lin = np.linspace(0.0, 1.0, s)
    gx, gy = np.meshgrid(lin, lin)
    return np.vstack([gx.ravel(), gy.ravel()]).T


def ball_offsets(s, r, ties_in=True):
    """Lattice offsets (dx,dy) with dx^2+dy^2 <= (r(s-1))^2.  TIE RULE: sklearn's
    expanded-form float64 distance makes offsets exactly on the sphere (e.g. (+-12,0) at s=241,
    r=0.05) land in or out by rounding; this generator uses exact integer arithmetic and includes
    them when ties_in (mathematical `<=`), excludes them otherwise."""
    rr = r * (s - 1)
    lim2 = rr * rr
    R = int(math.floor(rr + 1e-9))
    offs = []
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            d2 = dx * dx + dy * dy
            on_sphere = abs(d2 - lim2) <= 1e-9 * max(lim2, 1.0)
            if d2 < lim2 and not on_sphere:
                offs.append((dy, dx))
            elif on_sphere and ties_in:
                offs.append((dy, dx))
    return offs          # already sorted by (dy, dx) == ascending dst for a fixed src


def ball_connectivity(s, r, ties_in=True):
    """edge_index [2,E] int64, identical (incl. order) to
    np.vstack(np.where(pairwise_distances(grid) <= r)) (GNO/utilities.py:250-255) whenever no lattice
    distance ties with r.  Vectorised over offsets; E int64 pairs."""
    offs = ball_offsets(s, r, ties_in)
    iy, ix = np.meshgrid(np.arange(s), np.arange(s), indexing='ij')
    iy = iy.ravel()
    ix = ix.ravel()
    src_l, dst_l = [], []
    for (dy, dx) in offs:
        jy = iy + dy
        jx = ix + dx
        ok = (jy >= 0) & (jy < s) & (jx >= 0) & (jx < s)
        src_l.append((iy * s + ix)[ok])
        dst_l.append((jy * s + jx)[ok])
    src = np.concatenate(src_l)
    dst = np.concatenate(dst_l)
    order = np.lexsort((dst, src))
    return np.vstack([src[order], dst[order]]).astype(np.int64)


def ball_edge_attr(grid, edge_index, theta):
    """GNO/utilities.py:269-277 (f=None, theta given): [pos_src(2), pos_dst(2), theta_src, theta_dst]."""
    e = edge_index.shape[1]
    d = grid.shape[1]
    ea = np.zeros((e, 3 * d))
    ea[:, 0:2 * d] = grid[edge_index.T].reshape((e, -1))
    ea[:, 2 * d] = theta[edge_index[0]]
    ea[:, 2 * d + 1] = theta[edge_index[1]]
    return ea.astype(np.float32)


def reference_init(in_channels, out_channels, mlp_layers, root_weight=True, bias=True, seed=0):
    """Draw parameters exactly as the reference constructs them under torch.manual_seed(seed):
    DenseNet ctor (GNO/utilities.py:212-213: nn.Linear default init, in layer order), then
    NNConv_old.reset_parameters (GNO/nn_conv.py:261-265): reset(nn) re-inits every Linear in order,
    then uniform(in, root), uniform(in, bias) = U(+-1/sqrt(in))."""
    torch.manual_seed(seed)
    lins = [torch.nn.Linear(mlp_layers[j], mlp_layers[j + 1]) for j in range(len(mlp_layers) - 1)]
    root_t = torch.empty(in_channels, out_channels) if root_weight else None
    bias_t = torch.empty(out_channels) if bias else None
    for lin in lins:
        lin.reset_parameters()
    bound = 1.0 / math.sqrt(in_channels)
    if root_t is not None:
        root_t.uniform_(-bound, bound)
    if bias_t is not None:
        bias_t.uniform_(-bound, bound)
    return ([l.weight.detach().clone() for l in lins], [l.bias.detach().clone() for l in lins],
            root_t, bias_t)
