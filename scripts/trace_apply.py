"""Measurement aid: CTA timeline of one KernelNN conv stack (NNCONV_TRACE=1)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['NNCONV_TRACE'] = '1'
import numpy as np, torch
from graph_pde_b200 import _lib, graphs
from graph_pde_b200.models import KernelNN

wl = os.environ.get('WL', 'darcy241')
s, r = (241, 0.05) if wl == 'darcy241' else (85, 0.10)
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = KernelNN(64, 1024, 6, 6, in_width=6).to(dev).eval()
x, ei, ea = graphs.darcy_sample(s, r, dev, seed=0)
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
    model.conv1._h_cache.clear()
    model.conv_stack(x0, ei, ea); torch.cuda.synchronize()
rec = dump()
np.save('gpurun_out/trace_%s.npy' % wl, rec)
tag = (rec[:, 0] & 0xfff).astype(int); seq = (rec[:, 0] >> 12).astype(int)
t0 = rec[:, 3].astype(np.int64); t1 = rec[:, 4].astype(np.int64); t2 = rec[:, 5].astype(np.int64)
base = t0.min()
print('records', len(rec), 'span ms', (t2.max() - base) / 1e6)
# per launch summary, ordered by first start
keys = sorted(set(zip(tag, seq)), key=lambda k: t0[(tag == k[0]) & (seq == k[1])].min())
rows = []
for k in keys:
    m = (tag == k[0]) & (seq == k[1])
    rows.append((k[0], k[1], int(m.sum()), (t0[m].min() - base) / 1e3, (t0[m].max() - base) / 1e3, (t2[m].min() - base) / 1e3,
                 (t2[m].max() - base) / 1e3, float(np.mean(t2[m] - t0[m])) / 1e3, float(np.mean(t1[m] - t0[m])) / 1e3))
# print the chain of the 3rd apply (skip edge-feature kernels): find first 40 launches with tag 100/200 after the last 101
last_hidden = max(i for i, r_ in enumerate(rows) if r_[0] == 101)
print('tag seq ctas first_start last_start first_end last_end mean_dur mean_wait   (us)')
for r_ in rows[last_hidden + 1: last_hidden + 41]:
    print('%d %5d %4d %10.1f %10.1f %10.1f %10.1f %8.1f %8.1f' % r_)
# overlap statistics over the whole apply chain
chain = rows[last_hidden + 1:]
gaps = [chain[i + 1][3] - chain[i][6] for i in range(len(chain) - 1)]
print('next.first_start - prev.last_end (us): mean %.2f median %.2f min %.2f max %.2f' % (np.mean(gaps), np.median(gaps), np.min(gaps), np.max(gaps)))
