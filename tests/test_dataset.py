"""Data path: the host dataset classes against golden vectors produced by the reference's own
`CustomImageFolderDataset` / `InfiniteSampler` on tests/golden/slices.zip (oracle/gen_dataset_golden.py), the packed shard
round trip, and -- on the GPU -- the one-launch batch gather against the reference's `images.to(float32) / 127.5 - 1`."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle.gen_dataset_golden import SAMPLER_CASES, VARIANTS
from gan_track_b200.training import dataset as ds_mod

ZIP = os.path.join(ROOT, 'tests', 'golden', 'slices.zip')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'dataset.npz'))


def _open(name):
    return ds_mod.CustomImageFolderDataset(path=ZIP, dtype=np.float32, **VARIANTS[name])


@pytest.mark.parametrize('name', list(VARIANTS))
def test_zip_dataset_matches_reference_golden(gold, name):
    ds = _open(name)
    meta = json.loads(str(gold[f'{name}/meta']))
    assert len(ds) == meta['len'] and ds.image_shape == meta['image_shape'] and ds.label_shape == meta['label_shape']
    assert ds.label_dim == meta['label_dim'] and bool(ds.has_labels) == meta['has_labels'] and bool(ds.has_onehot_labels) == meta['has_onehot']
    assert ds.name == meta['name'] and ds.resolution == meta['resolution'] and ds.num_channels == meta['num_channels']
    for i in range(len(ds)):
        img, lab, fname = ds[i]
        assert img.dtype == np.float32 and img.flags['C_CONTIGUOUS']
        assert np.array_equal(img, gold[f'{name}/images'][i]) and np.array_equal(lab, gold[f'{name}/labels'][i])
        assert fname == str(gold[f'{name}/fnames'][i])
        d = ds.get_details(i)
        assert d.raw_idx == int(gold[f'{name}/raw_idx'][i]) and d.xflip == bool(gold[f'{name}/xflip'][i])
        assert np.array_equal(np.asarray(d.raw_label), gold[f'{name}/raw_label'][i])
    for sname, (rank, rep, seed) in SAMPLER_CASES.items():
        it = iter(ds_mod.InfiniteSampler(ds, rank=rank, num_replicas=rep, seed=seed))
        want = gold[f'{name}/sampler/{sname}']
        assert [int(next(it)) for _ in range(len(want))] == want.tolist()
    ds.close()


def test_zip_dataset_errors():
    with pytest.raises(IOError):
        ds_mod.CustomImageFolderDataset(path='/tmp/not_a_zip_dir', dtype=np.float32, **VARIANTS['train_2mod_labels'])
    with pytest.raises(IOError):
        ds_mod.CustomImageFolderDataset(path=ZIP, dtype=np.float32, **dict(VARIANTS['train_2mod_labels'], split='validation'))
    with pytest.raises(IOError):
        ds_mod.CustomImageFolderDataset(path=ZIP, resolution=32, dtype=np.float32, **VARIANTS['train_2mod_labels'])


@pytest.mark.parametrize('name', list(VARIANTS))
@pytest.mark.parametrize('pack', ['float32', 'float16', 'uint16'])
def test_packed_shard_round_trip(tmp_path, gold, name, pack):
    """write_packed + PackedShard reproduce the zip dataset: exactly for float32, to the stated quantisation otherwise, with the
    same max_size / xflip / label semantics (index arithmetic on the shard)."""
    kw = VARIANTS[name]
    src = _open(name)
    path = ds_mod.write_packed(src, str(tmp_path / f'{name}_{pack}.gtshard'), dtype=pack)
    sh = ds_mod.PackedShard(path, max_size=kw['max_size'], use_labels=kw['use_labels'], xflip=kw['xflip'], random_seed=kw['random_seed'])
    assert len(sh) == len(src) and sh.image_shape == src.image_shape and sh.label_shape == src.label_shape
    tol = {'float32': 0.0, 'float16': 0.0626, 'uint16': 0.5 / 257 + 1e-5}[pack]
    for i in range(len(sh)):
        img, lab, fname = sh[i]
        ref = gold[f'{name}/images'][i]
        assert np.abs(img - ref).max() <= tol
        if pack == 'float32':
            assert np.array_equal(img, ref)
        assert np.array_equal(lab, gold[f'{name}/labels'][i]) and fname == str(gold[f'{name}/fnames'][i])
    src.close()


def test_device_batcher_has_no_cpu_path(tmp_path):
    src = _open('train_2mod_labels')
    sh = ds_mod.PackedShard(ds_mod.write_packed(src, str(tmp_path / 's.gtshard')), use_labels=True)
    with pytest.raises(RuntimeError):
        ds_mod.DeviceBatcher(sh, 'cpu')
    with pytest.raises(AssertionError):
        ds_mod.PackedShard(ZIP)                               # not a packed shard
    src.close()


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(VARIANTS))
@pytest.mark.parametrize('pack', ['float32', 'float16', 'uint16'])
def test_device_batcher_matches_reference_batches(tmp_path, gold, name, pack):
    """One gather launch == the reference loop's batch: DataLoader items stacked, `.to(float32) / 127.5 - 1`, one-hot labels;
    bit-exact from a float32 shard, and equal to the same arithmetic on the decoded values for the compact formats."""
    kw = VARIANTS[name]
    src = _open(name)
    path = ds_mod.write_packed(src, str(tmp_path / f'{name}_{pack}.gtshard'), dtype=pack)
    sh = ds_mod.PackedShard(path, max_size=kw['max_size'], use_labels=kw['use_labels'], xflip=kw['xflip'], random_seed=kw['random_seed'])
    bat = ds_mod.DeviceBatcher(sh, 'cuda')
    order = gold[f'{name}/sampler/single'][:11]
    img, lab = bat.batch(order)
    assert img.dtype == torch.float32 and tuple(img.shape) == (len(order), *sh.image_shape)
    assert np.array_equal(lab.cpu().numpy(), gold[f'{name}/labels'][order])
    if pack == 'float32':
        assert np.array_equal(img.cpu().numpy(), gold[f'{name}/normalised'][order])
    want = torch.from_numpy(np.stack([sh[int(i)][0] for i in order])).to(torch.float32) / 127.5 - 1
    assert torch.equal(img.cpu(), want)
    # un-normalised delivery (what Trainer.train_step takes; it applies /127.5 - 1 itself, as the reference loop does)
    raw, _ = bat.batch(order, scale=1.0, shift=0.0)
    assert torch.equal(raw.cpu(), torch.from_numpy(np.stack([sh[int(i)][0] for i in order])))
    # device-side index tensors, the iterator, and loud failures
    img2, _ = bat.batch(torch.as_tensor(order).cuda())
    assert torch.equal(img2, img)
    it = bat.iterate(batch_size=4, rank=0, num_replicas=1, seed=5)
    a, _ = next(it)
    b, _ = next(it)
    assert torch.equal(torch.cat([a, b]).cpu(), torch.from_numpy(gold[f'{name}/normalised'][gold[f'{name}/sampler/single'][:8]])) or pack != 'float32'
    with pytest.raises(RuntimeError):
        bat.batch(np.array([len(sh)]))
    with pytest.raises(RuntimeError):
        ds_mod.DeviceBatcher(sh, 'cpu')
    src.close()


@pytest.mark.gpu
def test_trainer_fed_from_device_batcher_sees_reference_reals(tmp_path, gold):
    """Loader and loop connected: `DeviceBatcher.iterate()` already delivers `/127.5 - 1`-normalised slices, so they go into
    `Trainer.train_step(..., normalized=True)`; the images the loss receives must be exactly the reference loop's
    `real_img` (training_loop_mi_multimodal.py:313-318), not normalised a second time.  Device-side index tensors are range
    checked too (RuntimeError, not a device-side assert)."""
    from gan_track_b200.training import training_loop as tl
    name = 'train_1mod_flip_max6'
    kw = VARIANTS[name]
    src = _open(name)
    sh = ds_mod.PackedShard(ds_mod.write_packed(src, str(tmp_path / 'feed.gtshard')), max_size=kw['max_size'], use_labels=kw['use_labels'],
                            xflip=kw['xflip'], random_seed=kw['random_seed'])
    bat = ds_mod.DeviceBatcher(sh, 'cuda')
    c, h, w = sh.image_shape
    assert c == 1 and h == w
    cfg = tl.claro_config(resolution=h, batch=4, num_gpus=1, cbase=1024, cmax=64, map_depth=2, cond=bool(kw['use_labels']))
    trainer = tl.Trainer(cfg, rank=0, device='cuda')
    seen = []
    orig = trainer.loss.accumulate_gradients

    def spy(**k):
        seen.append(k['real_img'].detach().clone())
        return orig(**k)
    trainer.loss.accumulate_gradients = spy
    it = bat.iterate(batch_size=4, rank=0, num_replicas=1, seed=5)
    img, lab = next(it)
    if lab.shape[1] != cfg.common.c_dim:
        lab = torch.zeros([4, cfg.common.c_dim], device='cuda')
    trainer.train_step(img, lab, normalized=True)
    want = torch.from_numpy(gold[f'{name}/normalised'][gold[f'{name}/sampler/single'][:4]]).cuda()
    assert len(seen) >= 2 and all(torch.equal(s, want) for s in seen)
    assert float(want.min()) >= -1.0 and float(want.max()) <= 1.0 and float(want.max() - want.min()) > 0.5
    # the raw contract gives the same reals through the default path
    seen.clear()
    raw, _ = bat.batch(gold[f'{name}/sampler/single'][:4], scale=1.0, shift=0.0)
    trainer.train_step(raw, lab)
    assert all(torch.allclose(s, want, atol=1e-6) for s in seen)
    with pytest.raises(RuntimeError):
        bat.batch(torch.tensor([0, len(sh)], device='cuda'))
    with pytest.raises(RuntimeError):
        bat.batch(torch.tensor([-1], device='cuda'))
    src.close()
