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


def collate(items):
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
            setattr(out, k, torch.cat([v + o for v, o in zip(vals, offsets)], dim=1))
        else:
            setattr(out, k, torch.cat(vals, dim=0))
    out.batch = torch.cat([torch.full((it.num_nodes,), i, dtype=torch.long) for i, it in enumerate(items)])
    return out


class DataLoader(object):
    def __init__(self, dataset, batch_size=1, shuffle=False):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = list(range(len(self.dataset)))
        if self.shuffle:
            random.shuffle(order)
        for i in range(0, len(order), self.batch_size):
            yield collate([self.dataset[j] for j in order[i:i + self.batch_size]])
