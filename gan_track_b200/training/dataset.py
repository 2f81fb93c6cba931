"""Training-set access for the single-channel / multi-modality slice datasets Gan-track trains on.

Three layers:

* `CustomImageFolderDataset` -- reads the reference's on-disk format, a zip of per-slice pickles `{modality: HxW float}` plus
  `<split>/dataset.json` labels, with the same constructor, item contract `(image float32 CHW, label, fname)`, `max_size`,
  `xflip`, label handling and properties as S3/training/dataset_mi_multimodal.py:29-285 (S3 = /root/reference/src/models/
  stylegan3).  Host code; it is the parity anchor and the converter's input.
* `write_packed` / `PackedShard` -- one contiguous `[N,C,H,W]` array (float32 exact, or float16 / uint16 to halve it) + labels
  in a single file that is memory-mapped: no unzip, no unpickle, no per-item Python.
* `DeviceBatcher` -- the whole shard resident in HBM (CLARO / pelvis are a few GB; a B200 has 180); a training batch is ONE
  kernel (`gt_batch_gather`, csrc/batch_gather.cu) that gathers the sampled slices, applies the x-flip half of the doubled
  dataset, and the loop's `/127.5 - 1` normalisation (S3/training/training_loop_mi_multimodal.py:317), driven by an index
  vector from `InfiniteSampler` (S3/torch_utils/misc.py:111-142).  Nothing on the host touches pixels during training.
"""
import json
import os
import pickle
import zipfile

import numpy as np
import torch

from .. import _lib
from .. import dnnlib


def _index_tables(n_raw, max_size, xflip, random_seed):
    """(raw_idx int64, flip uint8): which stored slice each dataset item is and whether it is mirrored.  `max_size` keeps a seeded
    random subset in storage order; `xflip` appends a mirrored copy of the (sub)set (reference :57-68)."""
    raw = np.arange(n_raw, dtype=np.int64)
    if max_size is not None and n_raw > max_size:
        np.random.RandomState(random_seed).shuffle(raw)
        raw = np.sort(raw[:max_size])
    flip = np.zeros(raw.size, dtype=np.uint8)
    if xflip:
        raw, flip = np.tile(raw, 2), np.concatenate([flip, np.ones_like(flip)])
    return raw, flip


class Dataset(torch.utils.data.Dataset):
    """Item contract of the reference's base class: `ds[i] -> (image CHW, label, fname)`; subclasses provide `_load_raw_image(raw_idx)
    -> (CHW array, fname)` and `_load_raw_labels() -> [N] int64 | [N,L] float32 | None`."""

    def __init__(self, name, raw_shape, dtype, max_size=None, use_labels=False, xflip=False, split='train', modalities=None, random_seed=0):
        self._name, self._dtype, self._split = name, dtype, split
        self._modalities = ['MR_nonrigid_CT', 'MR_MR_T2'] if modalities is None else modalities
        self._raw_shape = list(raw_shape)
        self._use_labels = use_labels
        self._raw_labels = self._label_shape = None
        self._raw_idx, self._xflip = _index_tables(self._raw_shape[0], max_size, xflip, random_seed)

    # -- subclass hooks -------------------------------------------------------------------------------------------
    def _load_raw_image(self, raw_idx):
        raise NotImplementedError

    def _load_raw_labels(self):
        raise NotImplementedError

    def close(self):
        pass

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __getstate__(self):
        return dict(self.__dict__, _raw_labels=None)        # labels are re-read lazily in worker processes

    # -- labels ---------------------------------------------------------------------------------------------------
    def _get_raw_labels(self):
        if self._raw_labels is None:
            lab = self._load_raw_labels() if self._use_labels else None
            if lab is None:
                lab = np.zeros([self._raw_shape[0], 0], dtype=np.float32)
            assert isinstance(lab, np.ndarray) and lab.shape[0] == self._raw_shape[0] and lab.dtype in [np.float32, np.int64]
            if lab.dtype == np.int64:
                assert lab.ndim == 1 and np.all(lab >= 0)
            self._raw_labels = lab
        return self._raw_labels

    @property
    def has_onehot_labels(self):
        return self._get_raw_labels().dtype == np.int64

    @property
    def label_shape(self):
        if self._label_shape is None:
            lab = self._get_raw_labels()
            self._label_shape = [int(lab.max()) + 1] if lab.dtype == np.int64 else lab.shape[1:]
        return list(self._label_shape)

    @property
    def label_dim(self):
        shape = self.label_shape
        assert len(shape) == 1
        return shape[0]

    @property
    def has_labels(self):
        return any(d != 0 for d in self.label_shape)

    def get_label(self, idx):
        lab = self._get_raw_labels()[self._raw_idx[idx]]
        if lab.dtype != np.int64:
            return lab.copy()
        onehot = np.zeros(self.label_shape, dtype=np.float32)
        onehot[lab] = 1
        return onehot

    def get_details(self, idx):
        raw_idx = int(self._raw_idx[idx])
        return dnnlib.EasyDict(raw_idx=raw_idx, xflip=bool(self._xflip[idx]), raw_label=self._get_raw_labels()[raw_idx].copy())

    # -- items ----------------------------------------------------------------------------------------------------
    def __len__(self):
        return self._raw_idx.size

    def __getitem__(self, idx):
        image, fname = self._load_raw_image(self._raw_idx[idx])
        assert isinstance(image, np.ndarray) and list(image.shape) == self.image_shape and image.dtype == self._dtype
        if self._xflip[idx]:
            assert image.ndim == 3
            image = image[:, :, ::-1]                       # mirrored left-right
        return image.copy(), self.get_label(idx), fname      # .copy() makes the mirrored view contiguous

    # -- description ----------------------------------------------------------------------------------------------
    name = property(lambda self: self._name)
    dtype = property(lambda self: self._dtype)
    split = property(lambda self: self._split)
    modalities = property(lambda self: self._modalities)
    modatilies = modalities                                  # (sic) the reference's spelling of the same property, :144
    image_shape = property(lambda self: list(self._raw_shape[1:]))

    @property
    def num_channels(self):
        c, _, _ = self.image_shape
        return c

    @property
    def resolution(self):
        _, h, w = self.image_shape
        assert h == w
        return h


class CustomImageFolderDataset(Dataset):
    """Zip of `<split>/.../*.pickle` slices, each a dict {modality: HxW array}; channels = the requested modalities in order; labels
    from `<split>/dataset.json` = {"labels": [[path relative to <split>/, class | vector], ...]}."""

    def __init__(self, path, resolution=None, **super_kwargs):
        if os.path.splitext(path)[1].lower() != '.zip':
            raise IOError('Path must point to a directory or zip')
        self._path, self._type, self._zipfile = path, 'zip', None
        self._split, self._modalities = super_kwargs['split'], super_kwargs['modalities']
        self._all_fnames = set(self._zip().namelist())
        self._image_fnames = sorted(f for f in self._all_fnames if f.lower().endswith('.pickle') and self._split in f)
        if not self._image_fnames:
            raise IOError('No image files found in the specified path')
        probe, _ = self._load_raw_image(0)
        if resolution is not None and tuple(probe.shape[1:]) != (resolution, resolution):
            raise IOError('Image files do not match the specified resolution')
        super().__init__(name=os.path.splitext(os.path.basename(path))[0], raw_shape=[len(self._image_fnames), *probe.shape], **super_kwargs)

    def _zip(self):
        if self._zipfile is None:
            self._zipfile = zipfile.ZipFile(self._path)      # opened lazily so that every worker process gets its own handle
        return self._zipfile

    def close(self):
        z, self._zipfile = self._zipfile, None
        if z is not None:
            z.close()

    def __getstate__(self):
        return dict(super().__getstate__(), _zipfile=None)

    def _load_raw_image(self, raw_idx):
        fname = self._image_fnames[raw_idx]
        with self._zip().open(fname, 'r') as f:
            slices = pickle.load(f)
        assert len(self._modalities) > 0
        return np.stack([np.asarray(slices[m]).astype(np.float32) for m in self._modalities]), fname

    def _load_raw_labels(self):
        meta = f'{self._split}/dataset.json'
        if meta not in self._all_fnames:
            return None
        with self._zip().open(meta, 'r') as f:
            table = json.load(f)['labels']
        if table is None:
            return None
        table = dict(table)
        rows = np.array([table[os.path.relpath(f.replace('\\', '/'), f'{self._split}/')] for f in self._image_fnames])
        assert len(rows) == len(self._image_fnames)
        return rows.astype({1: np.int64, 2: np.float32}[rows.ndim])


# ---------------------------------------------------------------------------------------------------------------------
# packed shard
# ---------------------------------------------------------------------------------------------------------------------

_MAGIC = b'GTSHARD1'
_PACK_DTYPES = {'float32': np.float32, 'float16': np.float16, 'uint16': np.uint16}


def write_packed(dataset, path, dtype='float32'):
    """All raw slices of `dataset` (before max_size / xflip, which stay index arithmetic) as one `[N,C,H,W]` array + labels.
    float32 is exact; float16 rounds; uint16 stores round(x * 257) for data in [0, 255] (step 1/257).  Layout: magic, uint64
    header length, JSON header, padding to 4096, images, labels."""
    assert dtype in _PACK_DTYPES
    n = dataset._raw_shape[0]
    c, h, w = dataset.image_shape
    labels = dataset._get_raw_labels()
    header = dict(version=1, name=dataset.name, n=n, image_shape=[c, h, w], dtype=dtype, split=dataset.split, modalities=list(dataset.modalities),
                  label_dtype=str(labels.dtype), label_shape=list(labels.shape[1:]), fnames=list(getattr(dataset, '_image_fnames', [])))
    hj = json.dumps(header).encode()
    off = (len(_MAGIC) + 8 + len(hj) + 4095) // 4096 * 4096
    with open(path, 'wb') as f:
        f.write(_MAGIC)
        f.write(np.uint64(len(hj)).tobytes())
        f.write(hj)
        f.write(b'\0' * (off - f.tell()))
        for i in range(n):
            img, _ = dataset._load_raw_image(i)
            if dtype == 'uint16':
                img = np.clip(np.rint(img * 257.0), 0, 65535)
            f.write(np.ascontiguousarray(img.astype(_PACK_DTYPES[dtype])).tobytes())
        f.write(np.ascontiguousarray(labels).tobytes())
    return path


class PackedShard(Dataset):
    """Memory-mapped view of a `write_packed` file with the Dataset item contract (images decoded to float32)."""

    def __init__(self, path, max_size=None, use_labels=False, xflip=False, random_seed=0):
        with open(path, 'rb') as f:
            assert f.read(len(_MAGIC)) == _MAGIC, 'not a packed shard'
            hlen = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            self.header = json.loads(f.read(hlen).decode())
        h = self.header
        self._path = path
        off = (len(_MAGIC) + 8 + hlen + 4095) // 4096 * 4096
        n, (c, hh, ww) = h['n'], h['image_shape']
        self.pack_dtype = h['dtype']
        self.images = np.memmap(path, mode='r', dtype=_PACK_DTYPES[h['dtype']], offset=off, shape=(n, c, hh, ww))
        loff = off + self.images.nbytes
        lshape = (n, *h['label_shape'])
        self._packed_labels = np.array(np.memmap(path, mode='r', dtype=np.dtype(h['label_dtype']), offset=loff, shape=lshape)) if int(np.prod(lshape)) else None
        self._fnames = h.get('fnames') or [f'{i:08d}' for i in range(n)]
        super().__init__(name=h['name'], raw_shape=[n, c, hh, ww], dtype=np.float32, max_size=max_size, use_labels=use_labels, xflip=xflip,
                         split=h['split'], modalities=h['modalities'], random_seed=random_seed)

    def decode(self, raw):
        if self.pack_dtype == 'uint16':
            return raw.astype(np.float32) / np.float32(257.0)
        return raw.astype(np.float32)

    def _load_raw_image(self, raw_idx):
        return self.decode(np.asarray(self.images[raw_idx])), self._fnames[raw_idx]

    def _load_raw_labels(self):
        return self._packed_labels


# ---------------------------------------------------------------------------------------------------------------------
# device-resident batches
# ---------------------------------------------------------------------------------------------------------------------

class InfiniteSampler(torch.utils.data.Sampler):
    """Endless index stream with a sliding-window shuffle, strided over replicas (S3/torch_utils/misc.py:111-142)."""

    def __init__(self, dataset, rank=0, num_replicas=1, shuffle=True, seed=0, window_size=0.5):
        assert len(dataset) > 0 and num_replicas > 0 and 0 <= rank < num_replicas and 0 <= window_size <= 1
        self.dataset = dataset
        self.rank = rank
        self.num_replicas = num_replicas
        self.shuffle = shuffle
        self.seed = seed
        self.window_size = window_size

    def __iter__(self):
        order = np.arange(len(self.dataset))
        rnd = None
        window = 0
        if self.shuffle:
            rnd = np.random.RandomState(self.seed)
            rnd.shuffle(order)
            window = int(np.rint(order.size * self.window_size))
        idx = 0
        while True:
            i = idx % order.size
            if idx % self.num_replicas == self.rank:
                yield order[i]
            if window >= 2:
                j = (i - rnd.randint(window)) % order.size
                order[i], order[j] = order[j], order[i]
            idx += 1


_DT_CODE = {'float32': 0, 'float16': 1, 'uint16': 3}


class DeviceBatcher:
    """The shard's pixels and labels live in HBM; `batch(indices)` is one gather kernel producing what the reference loop
    builds from a DataLoader batch: `real_img = images.to(device).to(float32) / 127.5 - 1`, `real_c = labels.to(device)`
    (S3/training/training_loop_mi_multimodal.py:313-319).  `indices` are dataset indices (xflip doubling and max_size included,
    i.e. what `InfiniteSampler` yields)."""

    def __init__(self, shard, device):
        assert isinstance(shard, PackedShard)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('gan_track_b200: DeviceBatcher needs a CUDA device (this package has no CPU path)')
        self.shard = shard
        self.pack_dtype = shard.pack_dtype
        raw = np.array(shard.images)                 # one pass over the memory map; the copy is what gets uploaded
        if raw.dtype == np.uint16:
            raw = raw.view(np.int16)                 # torch has no uint16 arithmetic; the kernel reinterprets the bits
        self.images = torch.from_numpy(raw).to(self.device)
        self.raw_idx = torch.from_numpy(shard._raw_idx.copy()).to(self.device)
        self.xflip = torch.from_numpy(shard._xflip.copy()).to(self.device)
        self.labels = torch.from_numpy(np.stack([shard.get_label(i) for i in range(len(shard))])).to(self.device) if len(shard) else None
        self.sampler = None
        self.check_device_indices = True

    def batch(self, indices, scale=127.5, shift=-1.0):
        """-> (real_img float32 in [-1, 1] = raw / scale + shift, real_c).  NOTE the images are already normalised: pass them to
        `Trainer.train_step(..., normalized=True)`.  `scale=1, shift=0` returns the raw [0, 255] values of the loader contract."""
        lib = _lib.load()
        if not isinstance(indices, torch.Tensor) or not indices.is_cuda:
            host = np.asarray(indices.cpu() if isinstance(indices, torch.Tensor) else indices, dtype=np.int64)
            if host.size == 0 or host.min() < 0 or host.max() >= len(self.shard):
                raise RuntimeError(f'gan_track_b200: dataset index out of range [0, {len(self.shard)})')
            indices = torch.from_numpy(host)
        elif self.check_device_indices:
            # device-resident indices: one min / max reduction and a host read (switch off for a sync-free loop whose indices
            # are known good); without it a bad id would surface as NaN pixels plus a device-side assert in index_select
            lo, hi = (int(v) for v in torch.stack([indices.min(), indices.max()]).tolist()) if indices.numel() else (0, -1)
            if indices.numel() == 0 or lo < 0 or hi >= len(self.shard):
                raise RuntimeError(f'gan_track_b200: dataset index out of range [0, {len(self.shard)})')
        idx = indices.to(device=self.device, dtype=torch.int64, non_blocking=True)
        b = idx.numel()
        n, c, h, w = self.images.shape
        out = torch.empty([b, c, h, w], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.gt_batch_gather(_lib.ptr(self.images), _DT_CODE[self.pack_dtype], _lib.ptr(idx), _lib.ptr(self.raw_idx), _lib.ptr(self.xflip),
                                           _lib.ptr(out), b, c, h, w, n, int(self.raw_idx.numel()), float(scale), float(shift),
                                           _lib.stream_of(out)), 'gt_batch_gather')
        _lib.count_launch()
        return out, self.labels.index_select(0, idx)

    def iterate(self, batch_size, rank=0, num_replicas=1, seed=0):
        """Endless (real_img, real_c) batches in the order the reference's DataLoader + InfiniteSampler would deliver them."""
        it = iter(InfiniteSampler(self.shard, rank=rank, num_replicas=num_replicas, seed=seed))
        while True:
            yield self.batch(np.fromiter((next(it) for _ in range(batch_size)), dtype=np.int64, count=batch_size))
