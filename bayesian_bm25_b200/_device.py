"""Host <-> device marshalling helpers shared by the Python mirrors."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def dev_f64(x, device) -> torch.Tensor:
    """Contiguous float64 device tensor from array-like / tensor input."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(device)


def ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def elementwise(fn_name: str, arrays, extra_pre=(), extra_post=()):
    """Broadcast `arrays` (array-likes), run the named bb25 elementwise entry point
    on the current CUDA device, return (ndarray result, was_scalar)."""
    dev = _lib.require_cuda()
    arrs = [np.asarray(a, dtype=np.float64) for a in arrays]
    scalar = all(a.ndim == 0 for a in arrs)
    bc = np.broadcast_arrays(*arrs)
    shape = bc[0].shape
    d_in = [dev_f64(np.ascontiguousarray(a).ravel(), f"cuda:{dev}") for a in bc]
    n = int(d_in[0].numel())
    out = torch.empty(n, dtype=torch.float64, device=f"cuda:{dev}")
    fn = getattr(_lib.lib(), fn_name)
    _lib.check(fn(dev, *extra_pre, *[t.data_ptr() for t in d_in], *extra_post, n, out.data_ptr(),
                  _lib.stream_ptr()))
    res = out.cpu().numpy().reshape(shape)
    return res, scalar
