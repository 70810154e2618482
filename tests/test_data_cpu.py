"""Host input pipeline of scripts/train.py (munit_b200/data.py; reference scripts/data.py, utils.py:50-250,458-636)."""
import os

import numpy as np
import pytest
import torch

from munit_b200 import data as D


def _write_images(folder, n, size, seed):
    from PIL import Image

    os.makedirs(folder, exist_ok=True)
    rng = np.random.RandomState(seed)
    paths = []
    for i in range(n):
        p = os.path.join(folder, f"im{i:02d}.png")
        Image.fromarray(rng.randint(0, 255, (size[1], size[0], 3), dtype=np.uint8)).save(p)
        paths.append(p)
    return paths


def test_folder_and_list_loaders(tmp_path):
    root = str(tmp_path)
    for split in ("trainA", "trainB", "testA", "testB"):
        _write_images(os.path.join(root, split), 3, (90, 70), hash(split) % 1000)
    conf = dict(batch_size=2, num_workers=0, new_size=64, crop_image_height=48, crop_image_width=56, data_root=root)
    tra, trb, tea, teb = D.get_all_data_loaders(conf)
    x = next(iter(tra))
    assert x.shape == (2, 3, 48, 56) and x.dtype == torch.float32 and -1.0 <= float(x.min()) and float(x.max()) <= 1.0
    y = next(iter(tea))
    assert y.shape == (2, 3, 64, 64)  # test split: resize then a new_size x new_size crop (utils.py:88-97)
    # list-file variant reads the same files
    lst = os.path.join(root, "a.txt")
    with open(lst, "w") as f:
        f.write("\n".join(sorted(os.listdir(os.path.join(root, "trainA")))) + "\n\n")
    ld = D.get_data_loader_list(os.path.join(root, "trainA"), lst, 3, False, None, 70, 90, 0, True)
    z = next(iter(ld))
    folder = torch.stack([D.ImageFolder(os.path.join(root, "trainA"), D.image_transform(False, None, 70, 90, True))[i]
                          for i in range(3)])
    assert torch.equal(z, folder)
    with pytest.raises(RuntimeError):
        D.ImageFolder(os.path.join(root, "empty_does_not_exist"))


def test_paired_with_mask_shares_one_crop_and_flip(tmp_path):
    from PIL import Image

    root = str(tmp_path)
    a = _write_images(os.path.join(root, "a"), 2, (100, 80), 1)
    # b = a with a changed box; mask = white box on black
    b, m = [], []
    for i, p in enumerate(a):
        arr = np.array(Image.open(p))
        arr2 = arr.copy()
        arr2[20:50, 30:70] = 255 - arr2[20:50, 30:70]
        pb = os.path.join(root, f"b{i}.png")
        Image.fromarray(arr2).save(pb)
        mk = np.zeros(arr.shape[:2], np.uint8)
        mk[20:50, 30:70] = 255
        pm = os.path.join(root, f"m{i}.png")
        Image.fromarray(mk).save(pm)
        b.append(pb)
        m.append(pm)
    ds = D.PairedWithMask(a, b, m, None, 64, 72, train=True)
    for _ in range(4):
        xa, xb, mask = ds[0]
        assert xa.shape == xb.shape == (3, 64, 72) and mask.shape == (1, 64, 72)
        assert set(mask.unique().tolist()) <= {0.0, 1.0}
        differs = ((xa - xb).abs().sum(0, keepdim=True) > 0).float()
        # wherever the pair differs the mask is set: crop and flip were applied identically to all three
        assert float((differs * (1 - mask)).sum()) == 0.0
        assert float(differs.sum()) > 0


def test_seeded_construction_matches_reference_with_adaptation_heads():
    """Same seed -> the reference's initial weights bit for bit, including the domain classifiers that are built
    after the generator initialisation (trainer.py:162-179).  Needs /root/reference (build container only)."""
    import pytest
    import torch

    from oracle import munit_oracle as O
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference checkout not present")
    from munit_b200.trainer import MUNIT_Trainer

    cfg = O.config_256_core()
    cfg["adaptation"].update(adv_lambda=6, dfeat_lambda=1)
    torch.manual_seed(0)
    rt = ref_loader.make_trainer(cfg)
    torch.manual_seed(0)
    ot = MUNIT_Trainer(cfg)
    for name in ("domain_classifier_sr_a", "domain_classifier_sr_b", "gen", "dis_a"):
        rs, os_ = getattr(rt, name).state_dict(), getattr(ot, name).state_dict()
        assert list(rs.keys()) == list(os_.keys()), name
        for k in rs:
            assert torch.equal(rs[k], os_[k]), (name, k)
    assert torch.equal(rt.s_a, ot.s_a) and torch.equal(rt.s_b, ot.s_b)
