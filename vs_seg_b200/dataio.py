"""MONAI-free data plane for the VS_Seg entry points: host-side plumbing only.

The reference builds its input pipeline from MONAI 0.4.0 dict transforms + nibabel
(/root/reference/params/VSparams.py:169-335, :582-594), neither of which exists in this image.
This module provides just what `VSparams` needs, with the same names and semantics:
NIfTI-1 read/write, `LoadNiftid`, `AddChanneld`, `Orientationd("RAS")`, `NormalizeIntensityd`,
`SpatialPadd`, `RandFlipd`, `RandSpatialCropd`, `ToTensord`, `Compose`, `CacheDataset`,
`list_data_collate`, `NiftiSaver`, `set_determinism`, and a synthetic-case generator for the
no-dataset configurations (BASELINE configs 0 and 4).  Nothing here runs on the GPU hot path.
"""
from __future__ import annotations

import gzip
import os
import random
import struct

import numpy as np
import torch
from torch.utils.data import Dataset
from torch.utils.data.dataloader import default_collate

# ---- NIfTI-1 ---------------------------------------------------------------------------------------
_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


def read_nifti(path):
    """-> (array in file axis order [i,j,k,...], affine 4x4 float64)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError(f"{path}: not a NIfTI-1 file")
    endian = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(endian + "i", raw[:4])[0] != 348:
        raise ValueError(f"{path}: bad NIfTI-1 header size")
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype, = struct.unpack(endian + "h", raw[70:72])
    pixdim = struct.unpack(endian + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(endian + "3f", raw[108:120])
    qform_code, sform_code = struct.unpack(endian + "2h", raw[252:256])
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    if datatype not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype {datatype}")
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    off = int(vox_offset) if vox_offset >= 348 else 352
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and np.isfinite(slope):
            data = data.astype(np.float32) * slope + inter
    affine = np.eye(4)
    if sform_code > 0:
        affine[:3, :] = np.array(struct.unpack(endian + "12f", raw[280:328])).reshape(3, 4)
    elif qform_code > 0:
        b, c, d = struct.unpack(endian + "3f", raw[256:268])
        ox, oy, oz = struct.unpack(endian + "3f", raw[268:280])
        a = np.sqrt(max(0.0, 1.0 - b * b - c * c - d * d))
        R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                      [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                      [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine[:3, :3] = R * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac])
        affine[:3, 3] = (ox, oy, oz)
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return np.ascontiguousarray(data), affine


def write_nifti(path, data, affine=None):
    data = np.asarray(data)
    if data.dtype == np.bool_:
        data = data.astype(np.uint8)
    if data.dtype == np.int64:
        data = data.astype(np.int32)
    key = data.dtype.str[1:]
    if key not in _CODES:
        data = data.astype(np.float32)
        key = "f4"
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, _CODES[key])
    struct.pack_into("<h", hdr, 72, data.dtype.itemsize * 8)
    vox = np.sqrt((affine[:3, :3] ** 2).sum(0))
    struct.pack_into("<8f", hdr, 76, 1.0, *[float(v) for v in vox], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    struct.pack_into("<2h", hdr, 252, 0, 1)  # sform only
    struct.pack_into("<12f", hdr, 280, *[float(v) for v in affine[:3, :].reshape(-1)])
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + np.asfortranarray(data).astype(data.dtype.newbyteorder("<")).tobytes(order="F")
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(payload)


# ---- dict transforms -------------------------------------------------------------------------------
class Randomizable:
    def set_random_state(self, seed=None, state=None):
        self.R = state if state is not None else np.random.RandomState(seed)
        return self


class Compose(Randomizable):
    def __init__(self, transforms):
        self.transforms = list(transforms)
        self.set_random_state(seed=np.random.randint(2 ** 31))

    def set_random_state(self, seed=None, state=None):
        super().set_random_state(seed, state)
        for t in getattr(self, "transforms", []):
            if isinstance(t, Randomizable):
                t.set_random_state(seed=self.R.randint(2 ** 31))
        return self

    def first_random_index(self):
        for i, t in enumerate(self.transforms):
            if isinstance(t, Randomizable):
                return i
        return len(self.transforms)

    def __call__(self, data, start=0):
        for t in self.transforms[start:]:
            data = t(data)
        return data


class _Keyed:
    def __init__(self, keys):
        self.keys = [keys] if isinstance(keys, str) else list(keys)


class LoadNiftid(_Keyed):
    def __call__(self, d):
        d = dict(d)
        for k in self.keys:
            arr, aff = read_nifti(d[k])
            d[k + "_meta_dict"] = {"filename_or_obj": d[k], "affine": aff.copy(), "original_affine": aff.copy(),
                                   "spatial_shape": np.asarray(arr.shape)}
            d[k] = arr
        return d


class AddChanneld(_Keyed):
    def __call__(self, d):
        d = dict(d)
        for k in self.keys:
            d[k] = d[k][None]
        return d


class Orientationd(_Keyed):
    """Reorders/flips the spatial axes to the requested axis codes (only "RAS" is used, VSparams.py:212)."""

    def __init__(self, keys, axcodes="RAS"):
        super().__init__(keys)
        if axcodes != "RAS":
            raise NotImplementedError("only RAS orientation is implemented")

    def __call__(self, d):
        d = dict(d)
        for k in self.keys:
            meta = d.get(k + "_meta_dict")
            aff = np.asarray(meta["affine"]) if meta else np.eye(4)
            R = aff[:3, :3]
            perm = [int(np.argmax(np.abs(R[ax, :]))) for ax in range(3)]  # file axis that runs along world axis ax
            if sorted(perm) != [0, 1, 2]:
                continue  # oblique beyond repair: leave as is
            arr = np.transpose(d[k], [0] + [p + 1 for p in perm])
            new_aff = aff[:, perm + [3]].copy()
            for ax in range(3):
                if new_aff[ax, ax] < 0:
                    arr = np.flip(arr, axis=ax + 1)
                    n = arr.shape[ax + 1]
                    new_aff[:3, 3] = new_aff[:3, 3] + new_aff[:3, ax] * (n - 1)
                    new_aff[:3, ax] = -new_aff[:3, ax]
            d[k] = np.ascontiguousarray(arr)
            if meta:
                meta = dict(meta)
                meta["affine"] = new_aff
                d[k + "_meta_dict"] = meta
        return d


class NormalizeIntensityd(_Keyed):
    def __call__(self, d):
        d = dict(d)
        for k in self.keys:
            x = d[k].astype(np.float32)
            d[k] = (x - x.mean()) / x.std()
        return d


class SpatialPadd(_Keyed):
    """Symmetric constant pad up to spatial_size."""

    def __init__(self, keys, spatial_size):
        super().__init__(keys)
        self.spatial_size = list(spatial_size)

    def __call__(self, d):
        d = dict(d)
        for k in self.keys:
            x = d[k]
            pads = [(0, 0)]
            for have, want in zip(x.shape[1:], self.spatial_size):
                diff = max(want - have, 0)
                pads.append((diff // 2, diff - diff // 2))
            d[k] = np.pad(x, pads, mode="constant")
        return d


class RandFlipd(_Keyed, Randomizable):
    def __init__(self, keys, prob=0.1, spatial_axis=None):
        _Keyed.__init__(self, keys)
        self.prob, self.spatial_axis = prob, spatial_axis
        self.set_random_state(seed=0)

    def __call__(self, d):
        d = dict(d)
        if self.R.random_sample() < self.prob:
            for k in self.keys:
                d[k] = np.ascontiguousarray(np.flip(d[k], axis=self.spatial_axis + 1))
        return d


class RandSpatialCropd(_Keyed, Randomizable):
    def __init__(self, keys, roi_size, random_center=True, random_size=False):
        _Keyed.__init__(self, keys)
        self.roi_size, self.random_center = list(roi_size), random_center
        self.set_random_state(seed=0)

    def __call__(self, d):
        d = dict(d)
        shape = d[self.keys[0]].shape[1:]
        size = [min(r, s) for r, s in zip(self.roi_size, shape)]
        if self.random_center:
            start = [int(self.R.randint(0, s - r + 1)) for r, s in zip(size, shape)]
        else:
            start = [(s - r) // 2 for r, s in zip(size, shape)]
        sl = (slice(None),) + tuple(slice(a, a + r) for a, r in zip(start, size))
        for k in self.keys:
            d[k] = np.ascontiguousarray(d[k][sl])
        return d


class ToTensord(_Keyed):
    def __call__(self, d):
        d = dict(d)
        for k in self.keys:
            d[k] = torch.as_tensor(np.ascontiguousarray(d[k]).astype(np.float32))
        return d


# ---- datasets --------------------------------------------------------------------------------------
class ArrayDataset(Dataset):
    """`monai.data.Dataset`: applies the transform at access time."""

    def __init__(self, data, transform=None):
        self.data, self.transform = list(data), transform

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        return self.transform(self.data[i]) if self.transform else self.data[i]


class CacheDataset(ArrayDataset):
    """Caches the deterministic prefix of the transform chain (everything before the first random
    transform), applies the rest at access time — `monai.data.CacheDataset(cache_rate=1)` semantics."""

    def __init__(self, data, transform, cache_rate=1.0, num_workers=0):
        super().__init__(data, transform)
        self._first_random = transform.first_random_index()
        self._cache = [Compose(transform.transforms[:self._first_random])(item) for item in self.data]

    def __getitem__(self, i):
        return self.transform(self._cache[i], start=self._first_random)


def list_data_collate(batch):
    return default_collate(batch)


def set_determinism(seed=0):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


class NiftiSaver:
    """Writes a [C, X, Y, Z] array next to the source file's base name (MONAI NiftiSaver layout:
    <output_dir>/<base>/<base><postfix>.nii.gz)."""

    def __init__(self, output_dir="./", output_postfix="seg", output_ext=".nii.gz"):
        self.output_dir, self.output_postfix, self.output_ext = output_dir, output_postfix, output_ext

    def save(self, data, meta_data=None):
        """MONAI's NiftiSaver (reference VSparams.py:582-594) receives both `affine` (of the array after
        Orientationd) and `original_affine` (of the file on disk) and writes the data back in the ORIGINAL
        orientation.  Here the axis permutation and flips implied by the two affines are undone exactly; if they
        differ by more than that (a resampling would be needed) the array is written with its own `affine`."""
        arr = data.detach().cpu().numpy() if torch.is_tensor(data) else np.asarray(data)
        name = os.path.basename(str(meta_data["filename_or_obj"])) if meta_data else "output"
        for ext in (".nii.gz", ".nii"):
            if name.endswith(ext):
                name = name[:-len(ext)]
        post = f"_{self.output_postfix}" if self.output_postfix else ""
        path = os.path.join(self.output_dir, name, name + post + self.output_ext)
        arr = np.squeeze(arr, 0) if arr.ndim == 4 and arr.shape[0] == 1 else arr
        affine = None
        if meta_data:
            cur = meta_data.get("affine")
            orig = meta_data.get("original_affine")
            affine = np.asarray(cur if cur is not None else orig, dtype=np.float64)
            if cur is not None and orig is not None and arr.ndim == 3:
                arr, affine = _to_original_orientation(arr, np.asarray(cur, dtype=np.float64),
                                                       np.asarray(orig, dtype=np.float64))
        write_nifti(path, arr, affine)
        return path


def _to_original_orientation(arr, affine, original_affine, tol=1e-3):
    """Undo the axis permutation / flips between `affine` (describes arr) and `original_affine`.
    Returns (array in the original voxel order, the affine that describes it)."""
    T = np.linalg.inv(original_affine) @ affine   # current voxel index -> original voxel index
    M = T[:3, :3]
    src = [int(np.argmax(np.abs(M[o, :]))) for o in range(3)]   # current axis that runs along original axis o
    P = np.zeros((3, 3))
    for o, c in enumerate(src):
        P[o, c] = np.sign(M[o, c])
    if sorted(src) != [0, 1, 2] or np.abs(M - P).max() > tol:
        return arr, affine   # not a pure reorientation: keep the array as it is, described by its own affine
    out = np.transpose(arr, src)
    for o, c in enumerate(src):
        if M[o, c] < 0:
            out = np.flip(out, axis=o)
    return np.ascontiguousarray(out), original_affine


# ---- synthetic cases (no dataset can be downloaded here) --------------------------------------------
def synthetic_case(seed, shape):
    """N(0,1) noise + bright ellipsoid 'tumour'; label = ellipsoid mask."""
    g = np.random.RandomState(seed)
    x = g.standard_normal(shape).astype(np.float32)
    c = [(0.25 + 0.5 * g.random_sample()) * s for s in shape]
    r = [max(2.0, s / 9.0) for s in shape]
    ii, jj, kk = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    m = ((ii - c[0]) / r[0]) ** 2 + ((jj - c[1]) / r[1]) ** 2 + ((kk - c[2]) / r[2]) ** 2 <= 1.0
    return x + 2.0 * m, m.astype(np.uint8)


def make_synthetic_dataset(data_root, split_csv, dataset="T1", shape=(64, 64, 64)):
    """Writes <data_root>/input_data/<case>/vs_gk_{t1_refT1,seg_refT1}.nii.gz for every case of the split
    (file naming of VSparams.load_T1_or_T2_data, VSparams.py:176-181)."""
    import csv
    mod = "t1_refT1" if dataset == "T1" else "t2_refT2"
    seg = "seg_refT1" if dataset == "T1" else "seg_refT2"
    with open(split_csv) as f:
        cases = [row[0] for row in csv.reader(f) if row]
    for i, case in enumerate(cases):
        img, lab = synthetic_case(2000 + i, shape)
        aff = np.diag([0.4, 0.4, 1.5, 1.0])
        write_nifti(os.path.join(data_root, "input_data", case, f"vs_gk_{mod}.nii.gz"), img, aff)
        write_nifti(os.path.join(data_root, "input_data", case, f"vs_gk_{seg}.nii.gz"), lab, aff)
    return cases
