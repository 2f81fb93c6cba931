"""Input pipeline throughput: the reference's loader design -- zip of per-slice pickles read by a
torch DataLoader with 3 workers (REF/src/bash/claro-*.sh: --workers=3), pinned, then `.to(device).float() / 127.5 - 1` -- against
the device-resident packed shard + one gather launch per batch.  Synthetic 256x256 single-modality slices.
python tools/bench_loader.py [n_slices]"""
import json
import os
import pickle
import sys
import tempfile
import time
import zipfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gan_track_b200.training import dataset as ds_mod  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = 32
dev = torch.device('cuda', 0)
tmp = tempfile.mkdtemp()
zpath = os.path.join(tmp, 'claro_like.zip')
rng = np.random.RandomState(0)
labels = []
with zipfile.ZipFile(zpath, 'w', compression=zipfile.ZIP_STORED) as z:
    for i in range(N):
        fname = f'train/p{i // 64:03d}/s{i:05d}.pickle'
        z.writestr(fname, pickle.dumps({'CT': (rng.rand(256, 256) * 255).astype(np.float32)}, protocol=4))
        labels.append([os.path.relpath(fname, 'train/'), int(rng.randint(0, 2))])
    z.writestr('train/dataset.json', json.dumps({'labels': labels}))
kw = dict(split='train', modalities=['CT'], use_labels=True, xflip=True, max_size=None, random_seed=0)
ds = ds_mod.CustomImageFolderDataset(path=zpath, dtype=np.float32, **kw)

# (a) the reference's design (workers must open their own zip handle: the reference gets that from the spawn start method + __getstate__)
ds.close()
loader = iter(torch.utils.data.DataLoader(dataset=ds, sampler=ds_mod.InfiniteSampler(ds, seed=0), batch_size=B, pin_memory=True, num_workers=3,
                                          prefetch_factor=2))
for _ in range(4):
    next(loader)
t0 = time.perf_counter()
iters = 40
for _ in range(iters):
    img, c, _ = next(loader)
    x = img.to(dev).to(torch.float32) / 127.5 - 1
    c = c.to(dev)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'zip + pickle DataLoader (3 workers): {iters * B / dt:9.0f} img/s', flush=True)
del loader

# (b) packed shard resident in HBM
t0 = time.perf_counter()
spath = ds_mod.write_packed(ds, os.path.join(tmp, 'claro_like.gtshard'), dtype='float32')
print(f'one-time conversion to a packed shard: {time.perf_counter() - t0:.1f} s for {N} slices ({os.path.getsize(spath) / 1e6:.0f} MB)', flush=True)
sh = ds_mod.PackedShard(spath, use_labels=True, xflip=True)
bat = ds_mod.DeviceBatcher(sh, dev)
it = bat.iterate(batch_size=B, seed=0)
for _ in range(4):
    next(it)
torch.cuda.synchronize()
t0 = time.perf_counter()
iters = 400
for _ in range(iters):
    x, c = next(it)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'device-resident shard, 1 gather launch per batch (incl. host sampler): {iters * B / dt:9.0f} img/s', flush=True)
idx = torch.randint(0, len(sh), [B], device=dev)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(200):
    bat.batch(idx)
e.record()
e.synchronize()
us = s.elapsed_time(e) / 200 * 1e3
print(f'gather kernel + label select: {us:.1f} us per batch of {B} ({2 * B * 256 * 256 * 4 / us / 1e3:.0f} GB/s)', flush=True)
