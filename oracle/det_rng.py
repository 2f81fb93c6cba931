"""ORACLE tooling: a device-independent random stream.

torch's CPU and CUDA generators produce different numbers for the same seed, so a fixed-seed training run on the GPU cannot
be compared with the reference's `ref` path on the CPU.  Inside `deterministic_rng(seed)` the random entry points the
training iteration uses (torch.randn / rand / randn_like / rand_like and Tensor.random_) draw from ONE numpy stream on
the host and move the result to the requested device, so the reference on the CPU, the oracle-backed host code on the CPU
and the CUDA path all see the same z, noise, style-mixing cut-offs and path-length probes, in call order.
Test infrastructure only (tests/, oracle/gen_loss_curve.py)."""
import contextlib

import numpy as np
import torch


@contextlib.contextmanager
def deterministic_rng(seed):
    rs = np.random.RandomState(seed)
    saved = (torch.randn, torch.rand, torch.randn_like, torch.rand_like, torch.Tensor.random_)

    def _shape(args, kw):
        if 'size' in kw:
            return tuple(kw['size'])
        if len(args) == 1 and isinstance(args[0], (list, tuple, torch.Size)):
            return tuple(args[0])
        return tuple(int(a) for a in args)

    def _finish(a, kw, like=None):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dtype = kw.get('dtype', like.dtype if like is not None else torch.float32)
        device = kw.get('device', like.device if like is not None else 'cpu')
        return t.to(device=device, dtype=dtype)

    def randn(*args, **kw):
        return _finish(rs.standard_normal(_shape(args, kw)).astype(np.float32), kw)

    def rand(*args, **kw):
        return _finish(rs.random_sample(_shape(args, kw)).astype(np.float32), kw)

    def randn_like(x, **kw):
        return _finish(rs.standard_normal(tuple(x.shape)).astype(np.float32), kw, like=x)

    def rand_like(x, **kw):
        return _finish(rs.random_sample(tuple(x.shape)).astype(np.float32), kw, like=x)

    def random_(self, lo=0, hi=None, **kw):
        if hi is None:
            lo, hi = 0, lo
        v = rs.randint(int(lo), int(hi), size=tuple(self.shape))
        return self.copy_(torch.from_numpy(np.asarray(v)).to(self.dtype))

    torch.randn, torch.rand, torch.randn_like, torch.rand_like, torch.Tensor.random_ = randn, rand, randn_like, rand_like, random_
    try:
        yield rs
    finally:
        torch.randn, torch.rand, torch.randn_like, torch.rand_like, torch.Tensor.random_ = saved
