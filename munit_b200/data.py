"""Input pipeline for scripts/train.py (reference: scripts/data.py:9-154, scripts/utils.py:50-250, 458-636).
Default: decoding, resize, crop and flip on the host workers exactly as in the reference (PIL + torchvision
transforms, same order: flip -> resize -> random crop -> to-tensor -> normalise to [-1, 1]).  With `gpu_preproc: 1`
 the workers stop after flip -> resize and ship the uint8 HWC image (3 B per pixel instead of 12,
no float work on the host); crop + to-tensor + normalise run on the GPU (munit_u8_crop_normalize), bit-identical to
the host transforms.  Either way the trainer consumes the NCHW fp32 batches the reference consumes."""
from __future__ import annotations

import os
import random
from typing import Callable, List, Optional, Sequence

import torch
from torch.utils.data import DataLoader, Dataset

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp")


def default_loader(path: str):
    from PIL import Image

    return Image.open(path).convert("RGB")


def read_list(flist: str) -> List[str]:
    """One path per line (blank lines ignored) -- data.py:13-23."""
    with open(flist, "r") as f:
        return [ln.strip() for ln in f if ln.strip()]


def list_images(folder: str) -> List[str]:
    """Every image under `folder`, sorted, recursively -- data.py:113-123."""
    out = []
    for root, _, names in sorted(os.walk(folder)):
        out += [os.path.join(root, n) for n in sorted(names) if n.lower().endswith(IMG_EXTENSIONS)]
    return out


class ImageFilelist(Dataset):
    """data.py:26-49: `flist` is a list file (or a python list) of paths relative to `root`."""

    def __init__(self, root: str, flist, transform: Optional[Callable] = None, loader: Callable = default_loader):
        self.root, self.transform, self.loader = root, transform, loader
        self.imlist = read_list(flist) if isinstance(flist, str) else list(flist)

    def __getitem__(self, index):
        img = self.loader(os.path.join(self.root, self.imlist[index]))
        return self.transform(img) if self.transform is not None else img

    def __len__(self):
        return len(self.imlist)


class ImageFolder(Dataset):
    """data.py:126-154."""

    def __init__(self, root: str, transform: Optional[Callable] = None, return_paths: bool = False,
                 loader: Callable = default_loader):
        self.imgs = list_images(root)
        if not self.imgs:
            raise RuntimeError("Found 0 images in: " + root + "\nSupported image extensions are: " + ",".join(IMG_EXTENSIONS))
        self.root, self.transform, self.return_paths, self.loader = root, transform, return_paths, loader

    def __getitem__(self, index):
        path = self.imgs[index]
        img = self.loader(path)
        if self.transform is not None:
            img = self.transform(img)
        return (img, path) if self.return_paths else img

    def __len__(self):
        return len(self.imgs)


def image_transform(train: bool, new_size: Optional[int], height: int, width: int, crop: bool):
    """utils.py:218-240 / 706-728."""
    from torchvision import transforms

    steps = []
    if train:
        steps.append(transforms.RandomHorizontalFlip())
    if new_size is not None:
        steps.append(transforms.Resize(new_size))
    if crop:
        steps.append(transforms.RandomCrop((height, width)))
    steps += [transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))]
    return transforms.Compose(steps)


def _loader(dataset, batch_size, train, num_workers):
    return DataLoader(dataset=dataset, batch_size=batch_size, shuffle=train, drop_last=True, num_workers=num_workers,
                      pin_memory=True, persistent_workers=num_workers > 0)


def get_data_loader_list(root, file_list, batch_size, train, new_size=None, height=256, width=256, num_workers=4,
                         crop=True):
    return _loader(ImageFilelist(root, file_list, image_transform(train, new_size, height, width, crop)), batch_size,
                   train, num_workers)


def get_data_loader_folder(input_folder, batch_size, train, new_size=None, height=256, width=256, num_workers=4,
                           crop=True):
    return _loader(ImageFolder(input_folder, image_transform(train, new_size, height, width, crop)), batch_size, train,
                   num_workers)


def get_all_data_loaders(conf: dict):
    """train A, train B, test A, test B (utils.py:50-156): `data_root` with trainA/testA/trainB/testB folders, or
    per-split folder + list-file keys."""
    bs, nw = conf["batch_size"], conf["num_workers"]
    size_a = size_b = conf.get("new_size")
    if "new_size" not in conf:
        size_a, size_b = conf["new_size_a"], conf["new_size_b"]
    h, w = conf["crop_image_height"], conf["crop_image_width"]
    gpu = bool(conf.get("gpu_preproc", 0))
    folder_fn = get_gpu_data_loader_folder if gpu else get_data_loader_folder
    list_fn = get_gpu_data_loader_list if gpu else get_data_loader_list
    if "data_root" in conf:
        mk = lambda split, train, size: folder_fn(  # noqa: E731
            os.path.join(conf["data_root"], split), bs, train, size, h if train else size, w if train else size, nw, True)
        return mk("trainA", True, size_a), mk("trainB", True, size_b), mk("testA", False, size_a), mk("testB", False, size_b)
    mk = lambda dom, split, train, size: list_fn(  # noqa: E731
        conf[f"data_folder_{split}_{dom}"], conf[f"data_list_{split}_{dom}"], bs, train, size, h if train else size,
        w if train else size, nw, True)
    return mk("a", "train", True, size_a), mk("b", "train", True, size_b), mk("a", "test", False, size_a), mk("b", "test", False, size_b)


class PairedWithMask(Dataset):
    """Synthetic pairs (utils.py:458-581 without the segmentation maps): image A, image B and a binary mask that
    share ONE random flip / resize / crop, so identical pixels stay aligned (trainer.py:452-464 relies on it)."""

    def __init__(self, list_a: Sequence[str], list_b: Sequence[str], list_mask: Sequence[str], new_size, height, width,
                 train: bool = True, loader: Callable = default_loader):
        assert len(list_a) == len(list_b) == len(list_mask)
        self.a, self.b, self.m = list(list_a), list(list_b), list(list_mask)
        self.new_size, self.h, self.w, self.train, self.loader = new_size, height, width, train, loader

    def __len__(self):
        return len(self.a)

    def __getitem__(self, i):
        import torchvision.transforms.functional as TF
        from PIL import Image

        a, b = self.loader(self.a[i]), self.loader(self.b[i])
        m = Image.open(self.m[i]).convert("L")
        if self.train and random.random() > 0.5:
            a, b, m = TF.hflip(a), TF.hflip(b), TF.hflip(m)
        if self.new_size is not None:
            a, b = TF.resize(a, self.new_size), TF.resize(b, self.new_size)
            m = TF.resize(m, self.new_size, interpolation=TF.InterpolationMode.NEAREST)
        wd, ht = a.size
        top = random.randint(0, max(ht - self.h, 0)) if self.train else max(ht - self.h, 0) // 2
        left = random.randint(0, max(wd - self.w, 0)) if self.train else max(wd - self.w, 0) // 2
        a, b, m = (TF.crop(t, top, left, self.h, self.w) for t in (a, b, m))
        norm = lambda t: TF.normalize(TF.to_tensor(t), (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))  # noqa: E731
        return norm(a), norm(b), (TF.to_tensor(m) > 0.5).float()


def get_synthetic_data_loader(list_a, list_b, list_mask, batch_size, train, new_size=None, height=256, width=256,
                              num_workers=4):
    ds = PairedWithMask(read_list(list_a), read_list(list_b), read_list(list_mask), new_size, height, width, train)
    return _loader(ds, batch_size, train, num_workers)


# ---------------------------------------------------------------------------------------------------------------------
# GPU tail of the pipeline: host = decode + flip + resize (uint8), device = crop + ToTensor + Normalize
# ---------------------------------------------------------------------------------------------------------------------
class _U8WithCrop(Dataset):
    """Wraps an image dataset: item -> (uint8 HWC tensor of the flipped + resized image, top, left).  The random
    decisions are made by the very torchvision calls the host pipeline makes, in the same order (flip, then the crop
    offsets after the resize), so one worker seed gives the same sample either way."""

    def __init__(self, paths, root, train, new_size, height, width, crop, loader=default_loader):
        self.paths, self.root, self.train = list(paths), root, train
        self.new_size, self.h, self.w, self.crop, self.loader = new_size, height, width, crop, loader

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, i):
        import numpy as np
        from torchvision import transforms
        import torchvision.transforms.functional as TF

        img = self.loader(os.path.join(self.root, self.paths[i]) if self.root else self.paths[i])
        if self.train and torch.rand(1) < 0.5:  # transforms.RandomHorizontalFlip.forward
            img = TF.hflip(img)
        if self.new_size is not None:
            img = TF.resize(img, self.new_size)
        top = left = 0
        if self.crop:
            top, left, _, _ = transforms.RandomCrop.get_params(img, (self.h, self.w))
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()), top, left


def _collate_u8(items):
    return [it[0] for it in items], [it[1] for it in items], [it[2] for it in items]


class GpuPreprocLoader:
    """Iterates [B, 3, H, W] fp32 CUDA batches: uint8 images from the host workers, crop + normalise on the device."""

    def __init__(self, dataset: _U8WithCrop, batch_size, train, num_workers, device="cuda"):
        self.ds, self.bs, self.device = dataset, batch_size, torch.device(device)
        self.loader = DataLoader(dataset=dataset, batch_size=batch_size, shuffle=train, drop_last=True,
                                 num_workers=num_workers, pin_memory=True, persistent_workers=num_workers > 0,
                                 collate_fn=_collate_u8)

    def __len__(self):
        return len(self.loader)

    class _HostView:
        """loader.dataset[i] -> the fp32 sample, computed on the host (train.py picks its display images this way)."""

        def __init__(self, ds):
            self.ds = ds

        def __len__(self):
            return len(self.ds)

        def __getitem__(self, i):
            im, t, l = self.ds[i]
            h = self.ds.h if self.ds.crop else im.shape[0]
            w = self.ds.w if self.ds.crop else im.shape[1]
            x = im[t:t + h, l:l + w].permute(2, 0, 1).float().div(255)
            return (x - 0.5) / 0.5

    @property
    def dataset(self):
        return GpuPreprocLoader._HostView(self.ds)

    def __iter__(self):
        from . import kernels as K

        for imgs, tops, lefts in self.loader:
            h = self.ds.h if self.ds.crop else imgs[0].shape[0]
            w = self.ds.w if self.ds.crop else imgs[0].shape[1]
            out = torch.empty(len(imgs), 3, h, w, dtype=torch.float32, device=self.device)
            for i, (im, t, l) in enumerate(zip(imgs, tops, lefts)):
                K.u8_crop_normalize(im.to(self.device, non_blocking=True), t, l, False, out[i])
            yield out


def get_gpu_data_loader_folder(input_folder, batch_size, train, new_size=None, height=256, width=256, num_workers=4,
                               crop=True, device="cuda"):
    return GpuPreprocLoader(_U8WithCrop(list_images(input_folder), None, train, new_size, height, width, crop),
                            batch_size, train, num_workers, device)


def get_gpu_data_loader_list(root, file_list, batch_size, train, new_size=None, height=256, width=256, num_workers=4,
                             crop=True, device="cuda"):
    return GpuPreprocLoader(_U8WithCrop(read_list(file_list), root, train, new_size, height, width, crop), batch_size,
                            train, num_workers, device)
