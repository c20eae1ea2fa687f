"""Fused loss / normaliser epilogue of the reference training step:

    mse = F.mse_loss(out.view(-1, 1), batch.y.view(-1, 1))                     UAI1_full_resolution.py:263
    loss = torch.norm(out.view(-1) - batch.y.view(-1), 1); loss.backward()     :265-266
    l2 = myloss(u_normalizer.decode(out.view(B, -1)), u_normalizer.decode(batch.y.view(B, -1)))   :268

(LpLoss.rel: graph-neural-operator/utilities.py:184-199, UnitGaussianNormalizer.decode: :87-99) computed by ONE
CUDA pass (csrc/loss.cu) with the results left on the device -- the reference's two ``.item()`` host syncs per step
(:272-273) become one read per epoch.  ``loss`` is differentiable (its gradient, sign(out - y), comes out of the same
pass); mse and l2 are reporting values like in the reference.
"""
import ctypes

import torch

from . import _lib


class _FusedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, y, mean, std, eps, batch):
        L = _lib.lib()
        o32 = out.detach().contiguous().float().view(batch, -1)
        y32 = y.detach().contiguous().float().view(batch, -1)
        n = o32.size(1)
        res = torch.empty(4, device=out.device)
        ws = torch.empty(2 + 2 * batch, device=out.device)
        grad = torch.empty_like(o32)
        vp = ctypes.c_void_p
        with torch.cuda.device(out.device):
            _lib.check(L.nnconv_loss_epilogue(vp(o32.data_ptr()), vp(y32.data_ptr()),
                                              vp(mean.data_ptr()) if mean is not None else vp(0),
                                              vp(std.data_ptr()) if std is not None else vp(0), float(eps), batch, n, 1.0,
                                              vp(grad.data_ptr()), vp(res.data_ptr()), vp(ws.data_ptr()),
                                              vp(torch.cuda.current_stream(out.device).cuda_stream)))
        ctx.save_for_backward(grad)
        ctx.shape = out.shape
        ctx.mark_non_differentiable(res)
        return res[1].clone(), res

    @staticmethod
    def backward(ctx, g_loss, _g_res):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss).view(ctx.shape), None, None, None, None, None


def fused_losses(out, y, batch_size=1, normalizer=None):
    """Returns (loss, stats): ``loss`` = ||out - y||_1 (differentiable w.r.t. out), ``stats`` a device tensor
    [mse, l1, sum of per-sample relative L2 of the decoded fields, mean of the same].  ``normalizer``: an object with
    ``mean``, ``std``, ``eps`` like the reference's UnitGaussianNormalizer (None = identity decode)."""
    if not out.is_cuda:
        raise RuntimeError('graph_pde_b200.fused_losses: CUDA tensors only (there is no CPU path)')
    mean = std = None
    eps = 0.0
    if normalizer is not None:
        mean = normalizer.mean.detach().to(out.device).contiguous().float().view(-1)
        std = normalizer.std.detach().to(out.device).contiguous().float().view(-1)
        eps = float(normalizer.eps)
    return _FusedL1.apply(out, y, mean, std, eps, int(batch_size))
