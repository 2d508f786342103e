import torch


class StateSpacePath(object):
    def __init__(self, xs, ys):
        self._xs, self._ys = xs, ys

    @property
    def time_indexes(self):
        return torch.stack([x.time_index for x in self._xs])

    def get_paths(self):
        return torch.stack([x.value for x in self._xs], 0), torch.stack(self._ys, 0)
