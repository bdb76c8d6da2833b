"""tensorflow_probability stand-in (TEST INFRASTRUCTURE ONLY): only ``distributions.Normal(loc, scale).cdf``,
which TFP evaluates as ``ndtr((x - loc) / scale)`` [TF-sem]."""
import types

import torch as _t


class _Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def cdf(self, x):
        x = _t.as_tensor(x, dtype=_t.float32)
        return _t.special.ndtr((x - self.loc) / self.scale)


distributions = types.SimpleNamespace(Normal=_Normal)
