"""Measurement aid: per-batch flag waits inside the fused persistent application kernel (NNCONV_TRACE=1)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['NNCONV_TRACE'] = '1'
import numpy as np, torch
from graph_pde_b200 import _lib, graphs
from graph_pde_b200.models import KernelNN
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = KernelNN(64, 1024, 6, 6, in_width=6).to(dev).eval()
x, ei, ea = graphs.darcy_sample(241, 0.05, dev, seed=0)
L = _lib.lib()
def dump():
    cap = 1 << 20
    buf = (ctypes.c_ulonglong * (cap * 6))()
    n = ctypes.c_uint(0)
    _lib.check(L.nnconv_debug_trace_dump(buf, cap, ctypes.byref(n)))
    return np.frombuffer(buf, dtype=np.uint64)[: n.value * 6].reshape(-1, 6).copy()
with torch.no_grad():
    x0 = model.fc1(x)
    model.conv_stack(x0, ei, ea); torch.cuda.synchronize(); dump()
    out = model.conv1(x0, ei, ea); torch.cuda.synchronize()
rec = dump()
np.save('gpurun_out/trace_fused.npy', rec)
tag = (rec[:, 0] & 0xfff).astype(int); b = (rec[:, 0] >> 12).astype(int); cta = rec[:, 1].astype(int)
t0 = rec[:, 3].astype(np.int64); t1 = rec[:, 4].astype(np.int64); t2 = rec[:, 5].astype(np.int64)
k = rec[tag == 300]
print('kernel CTAs', len(k), 'duration us', (k[:, 5].astype(np.int64).max() - k[:, 3].astype(np.int64).min()) / 1e3)
for c in sorted(set(cta[tag == 301])):
    m = (tag == 301) & (cta == c)
    order = np.argsort(b[m]); w = (t1[m] - t0[m])[order] / 1e3; st = t0[m][order]
    per = np.diff(st) / 1e3
    print('CTA %3d conv okY wait: mean %.2f us, p50 %.2f, max %.2f | batch period mean %.2f us (n=%d)' % (c, w.mean(), np.median(w), w.max(), per.mean(), len(w)))
    m2 = (tag == 302) & (cta == c)
    order = np.argsort(b[m2]); w2 = (t1[m2] - t0[m2])[order] / 1e3; d2 = (t2[m2] - t1[m2])[order] / 1e3
    print('        Y okC wait : mean %.2f us, p50 %.2f, max %.2f | Y production after wait mean %.2f us' % (w2.mean(), np.median(w2), w2.max(), d2.mean()))
