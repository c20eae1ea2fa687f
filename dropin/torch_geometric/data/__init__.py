"""Data / DataLoader with PyG's collation rule restated: tensors are concatenated along dim 0 (dim 1 for keys
containing 'index'), and keys containing 'index' are offset by the cumulative node count -- which is why the
MGKN scripts pin batch_size = 1 (their *_range tensors would be offset too, neurips1_MGKN.py:137-139)."""
import random

import torch


class Data(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k in self.__dict__ if not k.startswith('_')]

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device, **kw):
        for k in self.keys:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, **kw))
        return self

    def cuda(self):
        return self.to('cuda')


def collate(items, _cache=None):
    if len(items) == 1:
        # a shallow copy: `batch.to(device)` (UAI1_full_resolution.py:259) must not move the dataset item in place
        return Data(**{k: getattr(items[0], k) for k in items[0].keys})
    out = Data()
    offset = 0
    offsets = []
    for it in items:
        offsets.append(offset)
        offset += it.num_nodes
    for k in items[0].keys:
        vals = [getattr(it, k) for it in items]
        if not torch.is_tensor(vals[0]):
            setattr(out, k, vals)
            continue
        if 'index' in k:
            # same mesh for every sample (the SAME tensor object, as the reference builds its datasets:
            # UAI1_full_resolution.py:128-157) -> the block-diagonal index of this batch size is built once and
            # reused, so the NNConv plan keyed on it is reused as well
            key = (k, id(vals[0]), vals[0]._version, len(vals), tuple(offsets))
            if _cache is not None and all(v is vals[0] for v in vals):
                hit = _cache.get(key)
                if hit is None:
                    hit = (vals[0], torch.cat([v + o for v, o in zip(vals, offsets)], dim=1))
                    _cache[key] = hit
                setattr(out, k, hit[1])
            else:
                setattr(out, k, torch.cat([v + o for v, o in zip(vals, offsets)], dim=1))
        else:
            setattr(out, k, torch.cat(vals, dim=0))
    out.batch = torch.cat([torch.full((it.num_nodes,), i, dtype=torch.long, device=items[0].x.device)
                           for i, it in enumerate(items)])
    return out


class DataLoader(object):
    """PyG DataLoader surface (dataset, batch_size, shuffle).  Extra keyword ``device``: the dataset is moved there
    ONCE (tensors shared between samples -- the mesh's edge_index -- stay shared), block-diagonal collation runs on
    that device, and same-mesh batches reuse one cached block-diagonal edge_index: the per-step
    ``batch.to(device)`` of the reference loops (UAI1_full_resolution.py:259) becomes a no-op and the NNConv plan /
    edge_index upload happens once."""

    def __init__(self, dataset, batch_size=1, shuffle=False, device=None):
        self.batch_size, self.shuffle = batch_size, shuffle
        self._cache = {}
        if device is not None:
            moved = {}

            def mv(t):
                hit = moved.get(id(t))
                if hit is None:
                    hit = (t, t.to(device))
                    moved[id(t)] = hit
                return hit[1]
            dataset = [Data(**{k: (mv(getattr(it, k)) if torch.is_tensor(getattr(it, k)) else getattr(it, k))
                               for k in it.keys}) for it in dataset]
        self.dataset = dataset

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = list(range(len(self.dataset)))
        if self.shuffle:
            random.shuffle(order)
        for i in range(0, len(order), self.batch_size):
            yield collate([self.dataset[j] for j in order[i:i + self.batch_size]], self._cache)
