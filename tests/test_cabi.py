"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/gantrack_b200.h declares.
No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'gantrack_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gt_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib_path():
    from gan_track_b200 import build
    return build.build()


def test_header_declares_something():
    syms = header_symbols()
    assert 'gt_bias_act' in syms and 'gt_upfirdn2d' in syms and len(syms) >= 6


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (gt_[a-z0-9_]+)', out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f'declared in the header but not exported: {missing}'


def test_library_loads_and_binding_covers_header(lib_path):
    import torch  # noqa: F401  (brings libcudart into the process, as the product binding does)
    from gan_track_b200 import _lib
    lib = _lib.load()
    assert lib.gt_abi_version() >= 1
    assert lib.gt_last_error() is not None
    assert sorted(_lib._PROTOTYPES) == header_symbols()


def test_sass_is_sm100a(lib_path):
    out = subprocess.run(['cuobjdump', '-lelf', lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out[:500]


def test_conv_kernels_are_tcgen05_tma(lib_path):
    """The convolution objects carry Blackwell tensor-core / TMA / TMEM SASS (UTCHMMA, UTMALDG, LDTM), not HMMA."""
    from gan_track_b200 import build
    for obj in ('conv_igemm.o', 'conv_wgrad.o'):
        sass = subprocess.run(['cuobjdump', '-sass', os.path.join(build.OBJ, obj)], capture_output=True, text=True).stdout
        assert 'UTCHMMA' in sass and 'UTMALDG' in sass and 'LDTM' in sass, obj
        assert 'HMMA.16' not in sass


def test_product_ops_refuse_cpu_tensors():
    import torch
    from gan_track_b200.torch_utils.ops import bias_act, conv2d_gradfix, upfirdn2d
    x = torch.randn(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        bias_act.bias_act(x, None, act='lrelu')
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(x, upfirdn2d.setup_filter([1, 3, 3, 1]))
    with pytest.raises(RuntimeError):
        conv2d_gradfix.conv2d(x, torch.randn(4, 4, 3, 3))
    with pytest.raises(NotImplementedError):
        bias_act.bias_act(x, None, impl='ref')


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gan_track_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(dirpath, f)
