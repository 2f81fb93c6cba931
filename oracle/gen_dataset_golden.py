"""Golden vectors for the data path, produced by the REAL reference classes in the build container
(TEST INFRASTRUCTURE: only tests may use what this writes).

Writes tests/golden/slices.zip -- a tiny synthetic dataset in the reference's on-disk format (zip of `<split>/<patient>/*.pickle`
slice dicts {modality: HxW float} + `<split>/dataset.json`, as produced by REF/src/data/dataset_tool_mi.py:754-880) -- and
tests/golden/dataset.npz: what S3/training/dataset_mi_multimodal.py:CustomImageFolderDataset returns for it under several
option sets, and the index streams of S3/torch_utils/misc.py:InfiniteSampler.

    python oracle/gen_dataset_golden.py
"""
import io
import json
import os
import pickle
import sys
import types
import zipfile

import numpy as np

S3 = '/root/reference/src/models/stylegan3'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

VARIANTS = {        # name -> constructor options (shared with tests/test_dataset.py)
    'train_2mod_labels': dict(split='train', modalities=['CT', 'MR'], use_labels=True, xflip=False, max_size=None, random_seed=0),
    'train_1mod_flip_max6': dict(split='train', modalities=['MR'], use_labels=True, xflip=True, max_size=6, random_seed=3),
    'test_2mod_nolabels': dict(split='test', modalities=['MR', 'CT'], use_labels=False, xflip=True, max_size=None, random_seed=0),
}
SAMPLER_CASES = {'single': (0, 1, 5), 'rank0of2': (0, 2, 7), 'rank1of2': (1, 2, 7)}     # name -> (rank, replicas, seed)


def build_zip(path):
    rng = np.random.RandomState(1234)
    labels = {'train': [], 'test': []}
    with zipfile.ZipFile(path, 'w', compression=zipfile.ZIP_STORED) as z:
        for split, n in [('train', 10), ('test', 4)]:
            for i in range(n):
                fname = f'{split}/patient{i % 3:02d}/slice_{i:04d}.pickle'
                d = {'CT': (rng.rand(16, 16) * 255.0), 'MR': (rng.rand(16, 16) * 255.0).astype(np.float32), 'unused': np.zeros((16, 16))}
                info = zipfile.ZipInfo(fname, date_time=(2020, 1, 1, 0, 0, 0))
                z.writestr(info, pickle.dumps(d, protocol=4))
                labels[split].append([os.path.relpath(fname, f'{split}/'), int(rng.randint(0, 3))])
            info = zipfile.ZipInfo(f'{split}/dataset.json', date_time=(2020, 1, 1, 0, 0, 0))
            z.writestr(info, json.dumps({'labels': labels[split]}))


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, S3)
    for m in ['matplotlib', 'matplotlib.pyplot', 'openpyxl']:
        sys.modules.setdefault(m, types.ModuleType(m))
    from torch_utils import misc
    from training import dataset_mi_multimodal as ref
    zpath = os.path.join(OUT, 'slices.zip')
    build_zip(zpath)
    out = {}
    for name, kw in VARIANTS.items():
        ds = ref.CustomImageFolderDataset(path=zpath, dtype=np.float32, **kw)
        items = [ds[i] for i in range(len(ds))]
        out[f'{name}/images'] = np.stack([it[0] for it in items])
        out[f'{name}/labels'] = np.stack([it[1] for it in items])
        out[f'{name}/fnames'] = np.array([it[2] for it in items])
        out[f'{name}/raw_idx'] = np.array([ds.get_details(i).raw_idx for i in range(len(ds))])
        out[f'{name}/xflip'] = np.array([ds.get_details(i).xflip for i in range(len(ds))])
        out[f'{name}/raw_label'] = np.stack([np.asarray(ds.get_details(i).raw_label) for i in range(len(ds))])
        out[f'{name}/meta'] = np.array(json.dumps(dict(len=len(ds), image_shape=ds.image_shape, label_shape=ds.label_shape, label_dim=ds.label_dim,
                                                       has_labels=bool(ds.has_labels), has_onehot=bool(ds.has_onehot_labels), name=ds.name,
                                                       resolution=ds.resolution, num_channels=ds.num_channels)))
        # what the training loop turns a batch into (S3/training/training_loop_mi_multimodal.py:317)
        import torch
        out[f'{name}/normalised'] = (torch.from_numpy(out[f'{name}/images']).to(torch.float32) / 127.5 - 1).numpy()
        for sname, (rank, rep, seed) in SAMPLER_CASES.items():
            # torch 2.11's Sampler.__init__ no longer takes the data source the reference passes (misc.py:117): set the fields
            # its __init__ would set and run the reference's own __iter__
            smp = misc.InfiniteSampler.__new__(misc.InfiniteSampler)
            smp.dataset, smp.rank, smp.num_replicas, smp.shuffle, smp.seed, smp.window_size = ds, rank, rep, True, seed, 0.5
            it = iter(smp)
            out[f'{name}/sampler/{sname}'] = np.array([int(next(it)) for _ in range(3 * len(ds) + 5)])
        ds.close()
    np.savez_compressed(os.path.join(OUT, 'dataset.npz'), **out)
    print('wrote', zpath, os.path.getsize(zpath), 'bytes;', 'dataset.npz', os.path.getsize(os.path.join(OUT, 'dataset.npz')), 'bytes')


if __name__ == '__main__':
    main()
