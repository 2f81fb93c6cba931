"""ORACLE tooling: compile the REFERENCE's own CUDA plugins for sm_100a into oracle/_ref/ (test infrastructure, never shipped).

The reference JIT-builds `bias_act_plugin` and `upfirdn2d_plugin` with torch.utils.cpp_extension at first use
(S3/torch_utils/custom_ops.py:59-155; sources OPS/bias_act.{cpp,cu,h}, OPS/upfirdn2d.{cpp,cu,h}; flags `--use_fast_math`,
OPS/bias_act.py:46, OPS/upfirdn2d.py:31).  The GPU box has no /root/reference, so the same translation units are compiled
HERE, from the sources where they lie (nothing is copied into the repository), with the reference's flags plus the sm_100a
target, into two Python extension modules:

    oracle/_ref/bias_act_plugin.so      exports bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)   (OPS/bias_act.cpp:94-97)
    oracle/_ref/upfirdn2d_plugin.so     exports upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)   (OPS/upfirdn2d.cpp:102-105)

`oracle/_ref/` is git-ignored but travels to the GPU box with the snapshot, like the product library.  Only tests/ and
tools/op_sweep.py import these modules (GPU-side parity oracle and the microbenchmark baseline of BASELINE.json config 5).
`python oracle/build_ref_plugin.py`; __graft_entry__.build() calls `build()` when /root/reference is present."""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
OPS = '/root/reference/src/models/stylegan3/torch_utils/ops'
PLUGINS = {
    'bias_act_plugin': ['bias_act.cpp', 'bias_act.cu'],
    'upfirdn2d_plugin': ['upfirdn2d.cpp', 'upfirdn2d.cu'],
}


def available():
    return all(os.path.exists(os.path.join(OPS, s)) for srcs in PLUGINS.values() for s in srcs)


def built():
    return all(os.path.exists(os.path.join(OUT, name + '.so')) for name in PLUGINS)


def build(verbose=False, force=False):
    if not available():
        raise RuntimeError(f'reference sources not found under {OPS}')
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    inc = []
    for p in ce.include_paths(True) + [sysconfig.get_paths()['include']]:
        inc += ['-isystem', p]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    import concurrent.futures

    def compile_one(job):
        name, p, obj = job
        cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '--use_fast_math', '--allow-unsupported-compiler',
               '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', f'-DTORCH_EXTENSION_NAME={name}', '-DTORCH_API_INCLUDE_EXTENSION_H',
               f'-D_GLIBCXX_USE_CXX11_ABI={abi}', '-D__CUDA_NO_HALF_OPERATORS__', '-D__CUDA_NO_HALF_CONVERSIONS__',
               '-D__CUDA_NO_BFLOAT16_CONVERSIONS__', '-D__CUDA_NO_HALF2_OPERATORS__', '-x', 'cu', *inc, '-c', p, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {p}:\n{(r.stdout + r.stderr)[-6000:]}')

    todo, jobs = [], []
    for name, srcs in PLUGINS.items():
        so = os.path.join(OUT, name + '.so')
        paths = [os.path.join(OPS, s) for s in srcs]
        if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(p) for p in paths + [__file__]):
            continue
        objs = [os.path.join(OUT, name + '_' + os.path.basename(p).replace('.', '_') + '.o') for p in paths]
        todo.append((name, so, objs))
        jobs += [(name, p, o) for p, o in zip(paths, objs)]
    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:          # the four translation units in parallel (~4 min)
        list(ex.map(compile_one, jobs))
    for name, so, objs in todo:
        cmd = [nvcc, '-shared', '-cudart', 'shared', '-o', so, *objs, '-L', torch_lib, '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda', '-ltorch',
               '-ltorch_python', '-Xlinker', '-rpath', '-Xlinker', torch_lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed for {name}:\n{r.stdout + r.stderr}')
        for o in objs:
            os.remove(o)
        if verbose:
            print(f'[oracle/_ref] built {so}')
    return OUT


def load():
    """-> (bias_act_plugin, upfirdn2d_plugin) modules; raises if they were not built (tests skip on that)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension modules)
    mods = []
    for name in PLUGINS:
        so = os.path.join(OUT, name + '.so')
        if not os.path.exists(so):
            raise FileNotFoundError(f'{so} is missing: run `python oracle/build_ref_plugin.py` in the build container')
        spec = importlib.util.spec_from_file_location(name, so)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


if __name__ == '__main__':
    build(verbose=True, force='--force' in sys.argv)
