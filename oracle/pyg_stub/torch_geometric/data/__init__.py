class Data(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if hasattr(v, 'to'):
                setattr(self, k, v.to(device))
        return self


class DataLoader(object):
    def __init__(self, dataset, batch_size=1, shuffle=False):
        assert batch_size == 1
        self.dataset = dataset

    def __iter__(self):
        return iter(self.dataset)

    def __len__(self):
        return len(self.dataset)
