"""MessagePassing restated (PyG 1.3-era semantics) -- see ../../__init__.py for scope."""
import inspect

import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target'):
        super().__init__()
        assert aggr in ('add', 'mean', 'max')
        assert flow in ('source_to_target', 'target_to_source')
        self.aggr = aggr
        self.flow = flow
        self.__message_args__ = inspect.getfullargspec(self.message)[0][1:]
        self.__update_args__ = inspect.getfullargspec(self.update)[0][2:]

    def propagate(self, edge_index, size=None, **kwargs):
        # flow source_to_target: j = edge_index[0] (source), i = edge_index[1] (target)
        i, j = (0, 1) if self.flow == 'target_to_source' else (1, 0)
        n = None
        margs = []
        for name in self.__message_args__:
            if name.endswith('_j') or name.endswith('_i'):
                t = kwargs[name[:-2]]
                n = t.size(0) if n is None else n
                idx = edge_index[j] if name.endswith('_j') else edge_index[i]
                margs.append(t.index_select(0, idx))
            else:
                margs.append(kwargs[name])
        if size is not None:
            n = size if isinstance(size, int) else size[i]
        msg = self.message(*margs)
        tgt = edge_index[i]
        if self.aggr in ('add', 'mean'):
            out = torch.zeros((n,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
            out.index_add_(0, tgt, msg)
            if self.aggr == 'mean':
                cnt = torch.zeros(n, dtype=msg.dtype, device=msg.device)
                cnt.index_add_(0, tgt, torch.ones_like(tgt, dtype=msg.dtype))
                out = out / cnt.clamp(min=1).view(-1, *([1] * (msg.dim() - 1)))
        else:
            out = torch.zeros((n,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
            out = out.scatter_reduce(0, tgt.view(-1, 1).expand_as(msg), msg, 'amax', include_self=False)
        uargs = [kwargs[name] for name in self.__update_args__]
        return self.update(out, *uargs)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
