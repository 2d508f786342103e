import torch


class TimeseriesState(dict):
    """State of a time series at one time index.  ``values`` may be a zero-argument callable (evaluated lazily,
    once, on first ``.value`` access) - the reference passes ``kernel.sample`` that way (proposals/linear.py:53)."""

    def __init__(self, time_index, values, event_shape: torch.Size):
        super().__init__()
        self.time_index = time_index if isinstance(time_index, torch.Tensor) else torch.tensor(time_index)
        self._values = values
        self.event_shape = event_shape

    @property
    def value(self) -> torch.Tensor:
        if callable(self._values):
            self._values = self._values()
        return self._values

    @value.setter
    def value(self, v):
        self._values = v

    @property
    def batch_shape(self):
        v = self.value
        return v.shape[: v.dim() - len(self.event_shape)]

    def copy(self, values):
        return TimeseriesState(self.time_index, values, self.event_shape)

    def propagate_from(self, values, time_increment=1):
        return TimeseriesState(self.time_index + time_increment, values, self.event_shape)
