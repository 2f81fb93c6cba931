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


class Dataset(torch.utils.data.Dataset):
    def __init__(self, name, raw_shape, dtype, max_size=None, use_labels=False, xflip=False, split='train', modalities=None, random_seed=0):
        self._name = name
        self._dtype = dtype
        self._split = split
        self._modalities = ['MR_nonrigid_CT', 'MR_MR_T2'] if modalities is None else modalities
        self._raw_shape = list(raw_shape)
        self._use_labels = use_labels
        self._raw_labels = None
        self._label_shape = None
        # max_size: a seeded random subset, kept in raw order; applied before the flip doubling (reference :59-62)
        self._raw_idx = np.arange(self._raw_shape[0], dtype=np.int64)
        if max_size is not None and self._raw_idx.size > max_size:
            np.random.RandomState(random_seed).shuffle(self._raw_idx)
            self._raw_idx = np.sort(self._raw_idx[:max_size])
        self._xflip = np.zeros(self._raw_idx.size, dtype=np.uint8)
        if xflip:
            self._raw_idx = np.tile(self._raw_idx, 2)
            self._xflip = np.concatenate([self._xflip, np.ones_like(self._xflip)])

    def _get_raw_labels(self):
        if self._raw_labels is None:
            self._raw_labels = self._load_raw_labels() if self._use_labels else None
            if self._raw_labels is None:
                self._raw_labels = np.zeros([self._raw_shape[0], 0], dtype=np.float32)
            assert isinstance(self._raw_labels, np.ndarray) and self._raw_labels.shape[0] == self._raw_shape[0]
            assert self._raw_labels.dtype in [np.float32, np.int64]
            if self._raw_labels.dtype == np.int64:
                assert self._raw_labels.ndim == 1 and np.all(self._raw_labels >= 0)
        return self._raw_labels

    def close(self):
        pass

    def _load_raw_image(self, raw_idx):
        raise NotImplementedError

    def _load_raw_labels(self):
        raise NotImplementedError

    def __getstate__(self):
        return dict(self.__dict__, _raw_labels=None)

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __len__(self):
        return self._raw_idx.size

    def __getitem__(self, idx):
        image, fname = self._load_raw_image(self._raw_idx[idx])
        assert isinstance(image, np.ndarray) and list(image.shape) == self.image_shape and image.dtype == self._dtype
        if self._xflip[idx]:
            assert image.ndim == 3
            image = image[:, :, ::-1]
        return image.copy(), self.get_label(idx), fname

    def get_label(self, idx):
        label = self._get_raw_labels()[self._raw_idx[idx]]
        if label.dtype == np.int64:
            onehot = np.zeros(self.label_shape, dtype=np.float32)
            onehot[label] = 1
            label = onehot
        return label.copy()

    def get_details(self, idx):
        d = dnnlib.EasyDict()
        d.raw_idx = int(self._raw_idx[idx])
        d.xflip = int(self._xflip[idx]) != 0
        d.raw_label = self._get_raw_labels()[d.raw_idx].copy()
        return d

    name = property(lambda self: self._name)
    dtype = property(lambda self: self._dtype)
    modatilies = property(lambda self: self._modalities)      # (sic) the reference's spelling, :144
    modalities = property(lambda self: self._modalities)
    split = property(lambda self: self._split)
    image_shape = property(lambda self: list(self._raw_shape[1:]))

    @property
    def num_channels(self):
        assert len(self.image_shape) == 3
        return self.image_shape[0]

    @property
    def resolution(self):
        assert len(self.image_shape) == 3 and self.image_shape[1] == self.image_shape[2]
        return self.image_shape[1]

    @property
    def label_shape(self):
        if self._label_shape is None:
            raw = self._get_raw_labels()
            self._label_shape = [int(np.max(raw)) + 1] if raw.dtype == np.int64 else raw.shape[1:]
        return list(self._label_shape)

    @property
    def label_dim(self):
        assert len(self.label_shape) == 1
        return self.label_shape[0]

    @property
    def has_labels(self):
        return any(x != 0 for x in self.label_shape)

    @property
    def has_onehot_labels(self):
        return self._get_raw_labels().dtype == np.int64


class CustomImageFolderDataset(Dataset):
    """Zip of `<split>/.../*.pickle` slices, each a dict {modality: HxW array}; channels = the requested modalities in order."""

    def __init__(self, path, resolution=None, **super_kwargs):
        self._path = path
        self._zipfile = None
        self._split = super_kwargs['split']
        self._modalities = super_kwargs['modalities']
        if self._file_ext(path) != '.zip':
            raise IOError('Path must point to a directory or zip')
        self._type = 'zip'
        self._all_fnames = set(self._get_zipfile().namelist())
        self._image_fnames = sorted(f for f in self._all_fnames if self._file_ext(f) == '.pickle' and self._split in f)
        if len(self._image_fnames) == 0:
            raise IOError('No image files found in the specified path')
        name = os.path.splitext(os.path.basename(path))[0]
        raw_shape = [len(self._image_fnames)] + list(self._load_raw_image(0)[0].shape)
        if resolution is not None and (raw_shape[2] != resolution or raw_shape[3] != resolution):
            raise IOError('Image files do not match the specified resolution')
        super().__init__(name=name, raw_shape=raw_shape, **super_kwargs)

    @staticmethod
    def _file_ext(fname):
        return os.path.splitext(fname)[1].lower()

    def _get_zipfile(self):
        if self._zipfile is None:
            self._zipfile = zipfile.ZipFile(self._path)
        return self._zipfile

    def _open_file(self, fname):
        return self._get_zipfile().open(fname, 'r')

    def close(self):
        try:
            if self._zipfile is not None:
                self._zipfile.close()
        finally:
            self._zipfile = None

    def __getstate__(self):
        return dict(super().__getstate__(), _zipfile=None)

    def _load_raw_image(self, raw_idx):
        fname = self._image_fnames[raw_idx]
        with self._open_file(fname) as f:
            p = pickle.load(f)
        assert len(self._modalities) > 0
        first = p[self._modalities[0]]
        out = np.zeros((len(self._modalities), first.shape[0], first.shape[1]), dtype=np.float32)
        for i, m in enumerate(self._modalities):
            out[i] = np.asarray(p[m]).astype('float32')
        return out, fname

    def _load_raw_labels(self):
        fname = f'{self._split}/dataset.json'
        if fname not in self._all_fnames:
            return None
        with self._open_file(fname) as f:
            labels = json.load(f)['labels']
        if labels is None:
            return None
        labels = dict(labels)
        labels = [labels[os.path.relpath(f.replace('\\', '/'), f'{self._split}/')] for f in self._image_fnames]
        assert len(labels) == len(self._image_fnames)
        labels = np.array(labels)
        return labels.astype({1: np.int64, 2: np.float32}[labels.ndim])


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

    def batch(self, indices, scale=127.5, shift=-1.0):
        lib = _lib.load()
        if not isinstance(indices, torch.Tensor) or not indices.is_cuda:
            host = np.asarray(indices.cpu() if isinstance(indices, torch.Tensor) else indices, dtype=np.int64)
            if host.size == 0 or host.min() < 0 or host.max() >= len(self.shard):
                raise RuntimeError(f'gan_track_b200: dataset index out of range [0, {len(self.shard)})')
            indices = torch.from_numpy(host)
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
