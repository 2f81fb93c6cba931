"""Parameters on flat buffers + fused Adam / EMA (csrc/optim.cu).

`FlatParams(module)` re-homes every parameter of a module into ONE flat fp32 buffer (each parameter becomes a contiguous view,
so `state_dict`, autograd and the networks are unaffected).  `FlatAdam` keeps the two Adam moments and a per-parameter step
count alongside and applies, per training phase, `flat /= num_gpus; nan_to_num; Adam` (the reference's
S3/training/training_loop_mi_multimodal.py:343-351 + `opt.step()`) as one kernel over the phase's flat gradient -- the
concatenation of the gradients of exactly those parameters that received one, which is static per phase; parameters without
a gradient are skipped like `torch.optim.Adam` skips `grad is None`.  `ema_update` is the G_ema lerp (:358-366) in one launch.
CUDA only; the CPU / eager paths of the trainer keep torch.optim.Adam.
"""
import numpy as np
import torch

from .. import _lib


class FlatParams:
    def __init__(self, module):
        self.params = list(module.parameters())
        assert self.params and all(p.dtype == torch.float32 and p.is_cuda for p in self.params)
        self.sizes = [p.numel() for p in self.params]
        # every parameter starts on a 256-byte boundary: kernels read parameters (biases, weights) with 16-byte vector loads and
        # bulk copies, which the views into this buffer must allow like separately allocated tensors do
        ALIGN = 64
        padded = [(n + ALIGN - 1) // ALIGN * ALIGN for n in self.sizes]
        self.offsets = np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)
        self.total = int(self.offsets[-1])
        self.flat = torch.zeros([self.total], dtype=torch.float32, device=self.params[0].device)
        with torch.no_grad():
            for p, off, n in zip(self.params, self.offsets[:-1], self.sizes):
                view = self.flat[int(off):int(off) + n].view(p.shape)
                view.copy_(p.data)                   # logical (row-major) order, whatever memory format the parameter had
                p.data = view


class FlatAdam:
    CHUNK = 4096

    def __init__(self, flat_params, lr, betas, eps):
        self.fp = flat_params
        self.lr, self.b1, self.b2, self.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        dev = flat_params.flat.device
        self.m = torch.zeros_like(flat_params.flat)
        self.v = torch.zeros_like(flat_params.flat)
        self.steps = torch.zeros([len(flat_params.params)], dtype=torch.float32, device=dev)
        self._tables = {}

    def _table(self, key, active_idx):
        """(chunk records, active flags) for one phase: chunks of <= CHUNK elements that never straddle a parameter, with the
        parameter's offset in the flat parameter buffer and in the phase's compact gradient buffer."""
        key = (key, tuple(active_idx))
        if key in self._tables:
            return self._tables[key]
        lib = _lib.load()
        assert lib.gt_adam_chunk_bytes() == 24
        rec = np.dtype([('pstart', '<i8'), ('gstart', '<i8'), ('count', '<i4'), ('seg', '<i4')])
        rows = []
        goff = 0
        for i in active_idx:
            n, poff = self.fp.sizes[i], int(self.fp.offsets[i])
            for s in range(0, n, self.CHUNK):
                rows.append((poff + s, goff + s, min(self.CHUNK, n - s), i))
            goff += n
        arr = np.array(rows, dtype=rec)
        dev = self.fp.flat.device
        chunks = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        active = torch.zeros([len(self.fp.params)], dtype=torch.int32)
        active[list(active_idx)] = 1
        self._tables[key] = (chunks, len(rows), active.to(dev), goff)
        return self._tables[key]

    def step(self, key, active_idx, flat_grad, grad_scale=1.0):
        """flat_grad: fp32 [sum of the active parameters' sizes], in parameter order (what torch.cat of their gradients gives)."""
        chunks, nchunks, active, gtotal = self._table(key, tuple(active_idx))
        assert flat_grad.dtype == torch.float32 and flat_grad.is_contiguous() and flat_grad.numel() == gtotal
        lib = _lib.load()
        with torch.cuda.device(flat_grad.device):
            _lib.check(lib.gt_adam_flat(_lib.ptr(self.fp.flat), _lib.ptr(flat_grad), _lib.ptr(self.m), _lib.ptr(self.v), _lib.ptr(self.steps),
                                        _lib.ptr(active), len(self.fp.params), _lib.ptr(chunks), nchunks, self.lr, self.b1, self.b2, self.eps,
                                        float(grad_scale), 1e5, -1e5, _lib.stream_of(flat_grad)), 'gt_adam_flat')
        _lib.count_launch(2)


def ema_update(ema_flat, src_flat, weight):
    """ema += weight * (src - ema) over whole flat parameter buffers (torch._foreach_lerp_(ema_params, src_params, weight))."""
    assert ema_flat.flat.numel() == src_flat.flat.numel()
    lib = _lib.load()
    with torch.cuda.device(ema_flat.flat.device):
        _lib.check(lib.gt_ema_flat(_lib.ptr(ema_flat.flat), _lib.ptr(src_flat.flat), ema_flat.flat.numel(), float(weight),
                                   _lib.stream_of(ema_flat.flat)), 'gt_ema_flat')
    _lib.count_launch()
