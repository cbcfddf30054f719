"""API-compatible degradation operators (reference testing/operators/reverb.py:8-87).

The samplers also accept the reference's own operator objects (they only read `.params`); these classes exist so the
path can be used and tested without the reference on sys.path."""
import torch

from .spectral import LossSTFT, RirConv


class RIROperator:
    """Informed operator: convolution with a known room impulse response."""

    def __init__(self, op_hp=None, time_kernel_size=10, sample_rate=16000):
        self.time_kernel_size = time_kernel_size
        self.sample_rate = sample_rate
        self.params = None
        self._conv = None
        self._stft = None

    def update_params(self, k, **ignored):
        self.params = torch.as_tensor(k)
        self._conv = None

    def degradation(self, x, rm_delay=False, **ignored):
        assert self.params is not None, "filter is None"
        if rm_delay:
            raise NotImplementedError("rm_delay is not used by the samplers")
        squeeze = x.dim() == 1
        x2 = (x[None] if squeeze else x).float().contiguous()
        if self._conv is None or self._conv.n != x2.shape[1]:
            self._conv = RirConv(self.params.to(x2.device), x2.shape[1], x2.device)
        y = self._conv.forward(x2)
        return y[0] if squeeze else y

    def apply_stft(self, x):
        x2 = (x[None] if x.dim() == 1 else x).float().contiguous()
        if self._stft is None:
            self._stft = LossSTFT(x2.device)
        return torch.view_as_complex(self._stft.forward(x2))

    def get_time_RIR(self):
        return self.params
