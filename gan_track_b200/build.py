"""In-tree build of the C-ABI library `libgantrack_b200.so` (sm_100a only).

`python -m gan_track_b200.build` or `__graft_entry__.build()`.  Each `csrc/*.cu` is compiled to an object with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` (cross-compiles without a GPU) and the objects are linked
into one shared library next to this file, so the built `.so` travels with the repo snapshot to the GPU box.
The library links the SHARED libcudart (the same libcudart.so.12 torch has already loaded, so device/stream state
is shared with the host process); it does not link against torch.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libgantrack_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v',
]


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libgantrack_b200.so')
    return nvcc


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src, *extra))


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers += [os.path.join(HERE, '..', 'include', 'gantrack_b200.h')]
    headers = [h for h in headers if os.path.exists(h)]
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    jobs = []
    for f in sources:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-3] + '.o')
        if force or _newer(src, obj, headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, '-I', CSRC, '-I', os.path.join(HERE, '..', 'include'), '-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + '.log', 'w') as fh:
            fh.write(log)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{log[-8000:]}')
        return src, log

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f'[build] compiled {os.path.basename(src)}')
                    spills = [ln for ln in log.splitlines() if 'spill' in ln and '0 bytes spill stores, 0 bytes spill loads' not in ln]
                    for ln in spills:
                        print('   ', ln.strip())
    objs = [os.path.join(OBJ, f[:-3] + '.o') for f in sources]
    if jobs or force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc, '-shared', '-cudart', 'shared', '-o', LIB, *objs, '-ldl',
               '-Xlinker', '-rpath', '-Xlinker', '/usr/local/cuda/lib64']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
        if verbose:
            print(f'[build] linked {LIB}')
    return LIB


if __name__ == '__main__':
    build(verbose=True, force='--force' in sys.argv)
