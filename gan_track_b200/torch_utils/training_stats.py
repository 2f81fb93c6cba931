"""Running statistics with one all-reduce per `Collector.update()` (interface of S3/torch_utils/training_stats.py:
`init_multiprocessing`, `report`, `report0`, `Collector` with `update`, `mean`, `std`, `num`, `as_dict`, `[name]`).

Layout difference from the reference: all statistics of a device live in ONE persistent float64 table
`[max_names, 3]` (count, sum, sum of squares); `report()` adds one row with a single fused `index_add_`-free
expression (two reductions + one add into a view) and never allocates after the first call for a name, and
`_sync()` all-reduces the used part of the table in place -- the same single collective the reference issues
(:254-256) without re-stacking per-name tensors.
"""
import re

import numpy as np
import torch

_num_moments = 3
_counter_dtype = torch.float64
_max_names = 256
_rank = 0
_sync_device = None
_sync_called = False
_tables = dict()        # device -> float64 [max_names, 3]
_name_to_row = dict()   # name -> row index (same order on every rank)
_cumulative = dict()    # name -> float64 [3] on the CPU


def init_multiprocessing(rank, sync_device):
    global _rank, _sync_device
    assert not _sync_called
    _rank = rank
    _sync_device = sync_device


def _row(name):
    if name not in _name_to_row:
        assert len(_name_to_row) < _max_names
        _name_to_row[name] = len(_name_to_row)
    return _name_to_row[name]


def _table(device):
    t = _tables.get(device)
    if t is None:
        t = torch.zeros([_max_names, _num_moments], dtype=_counter_dtype, device=device)
        _tables[device] = t
    return t


def report(name, value):
    row = _row(name)
    elems = torch.as_tensor(value)
    if elems.numel() == 0:
        return value
    elems = elems.detach().flatten().to(torch.float32)
    moments = torch.stack([elems.new_full([], float(elems.numel())), elems.sum(), elems.square().sum()])
    _table(elems.device)[row].add_(moments.to(_counter_dtype))
    return value


def report0(name, value):
    report(name, value if _rank == 0 else [])
    return value


def _sync(names):
    if len(names) == 0:
        return []
    global _sync_called
    _sync_called = True
    device = _sync_device if _sync_device is not None else torch.device('cpu')
    used = len(_name_to_row)
    total = torch.zeros([used, _num_moments], dtype=_counter_dtype, device=device)
    for table in _tables.values():
        total.add_(table[:used].to(device))
        table[:used].zero_()
    if _sync_device is not None:
        torch.distributed.all_reduce(total)
    total = total.cpu()
    for name, row in _name_to_row.items():
        if name not in _cumulative:
            _cumulative[name] = torch.zeros([_num_moments], dtype=_counter_dtype)
        _cumulative[name].add_(total[row])
    return [(name, _cumulative[name]) for name in names]


class Collector:
    def __init__(self, regex='.*', keep_previous=True):
        self._regex = re.compile(regex)
        self._keep_previous = keep_previous
        self._cumulative = dict()
        self._moments = dict()
        self.update()
        self._moments.clear()

    def names(self):
        return [name for name in _name_to_row if self._regex.fullmatch(name)]

    def update(self):
        if not self._keep_previous:
            self._moments.clear()
        for name, cumulative in _sync(self.names()):
            if name not in self._cumulative:
                self._cumulative[name] = torch.zeros([_num_moments], dtype=_counter_dtype)
            delta = cumulative - self._cumulative[name]
            self._cumulative[name].copy_(cumulative)
            if float(delta[0]) != 0:
                self._moments[name] = delta

    def _get_delta(self, name):
        assert self._regex.fullmatch(name)
        if name not in self._moments:
            self._moments[name] = torch.zeros([_num_moments], dtype=_counter_dtype)
        return self._moments[name]

    def num(self, name):
        return int(self._get_delta(name)[0])

    def mean(self, name):
        delta = self._get_delta(name)
        if int(delta[0]) == 0:
            return float('nan')
        return float(delta[1] / delta[0])

    def std(self, name):
        delta = self._get_delta(name)
        if int(delta[0]) == 0 or not np.isfinite(float(delta[1])):
            return float('nan')
        if int(delta[0]) == 1:
            return float(0)
        mean = float(delta[1] / delta[0])
        raw_var = float(delta[2] / delta[0])
        return np.sqrt(max(raw_var - np.square(mean), 0))

    def as_dict(self):
        from ..dnnlib import EasyDict
        stats = EasyDict()
        for name in self.names():
            stats[name] = EasyDict(num=self.num(name), mean=self.mean(name), std=self.std(name))
        return stats

    def __getitem__(self, name):
        return self.mean(name)
