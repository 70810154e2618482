"""scripts/train.py end to end on a tiny synthetic dataset: loaders -> dis/gen updates with the reference's
ratio_disc_gen schedule -> sample grids -> checkpoints -> --resume; once plain, once with masks (recon_mask: 1)
and synthetic pairs (synthetic_frequency: 1, recon_synth_w > 0)."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import munit_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _images(folder, n, seed, size=(80, 72)):
    from PIL import Image

    os.makedirs(folder, exist_ok=True)
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        p = os.path.join(folder, f"im{i:02d}.png")
        Image.fromarray(rng.randint(0, 255, (size[1], size[0], 3), dtype=np.uint8)).save(p)
        out.append(p)
    return out


def _train_main():
    spec = importlib.util.spec_from_file_location("munit_train_script", os.path.join(ROOT, "scripts", "train.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.main


def _config(tmp, **over):
    cfg = O.config_256_core(crop_image_height=64, crop_image_width=64)
    cfg.update(new_size=64, batch_size=1, num_workers=0, display_size=2, max_iter=3, image_save_iter=2,
               image_display_iter=1, snapshot_save_iter=2, ratio_disc_gen=2, data_root=os.path.join(tmp, "data"))
    cfg.update(over)
    path = os.path.join(tmp, "tiny.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path


def test_train_script_runs_and_resumes(tmp_path):
    tmp = str(tmp_path)
    for k, split in enumerate(("trainA", "trainB", "testA", "testB")):
        _images(os.path.join(tmp, "data", split), 3, k)
    main = _train_main()
    cfg_path = _config(tmp)
    assert main(["--config", cfg_path, "--output_path", tmp]) == 3
    out = os.path.join(tmp, "outputs", "tiny")
    for f in ("config.yaml", "images/gen_a2b_train_00000002.jpg", "images/gen_b2a_test_00000002.jpg",
              "images/gen_a2b_train_current.jpg", "checkpoints/gen_00000002.pt", "checkpoints/dis_00000002.pt",
              "checkpoints/optimizer.pt"):
        assert os.path.isfile(os.path.join(out, f)), f
    sd = torch.load(os.path.join(out, "checkpoints", "gen_00000002.pt"), map_location="cpu")
    assert set(sd.keys()) == {"2"}  # shared-style generator checkpoint layout (trainer.py:1412)
    # resume from iteration 2 and run to 5
    cfg_path = _config(tmp, max_iter=5)
    assert main(["--config", cfg_path, "--output_path", tmp, "--resume"]) == 5
    assert os.path.isfile(os.path.join(out, "checkpoints", "gen_00000004.pt"))


def test_train_script_masks_and_synthetic_pairs(tmp_path):
    from PIL import Image

    tmp = str(tmp_path)
    lists = {}
    for k, split in enumerate(("trainA", "trainB", "testA", "testB")):
        paths = _images(os.path.join(tmp, "data", split), 3, 10 + k)
        lists[split] = paths
    # masks aligned with the train lists, synthetic pairs (b = a outside a box) + their masks
    def write_list(name, items):
        p = os.path.join(tmp, name)
        with open(p, "w") as f:
            f.write("\n".join(items) + "\n")
        return p

    masks = {}
    for dom, split in (("a", "trainA"), ("b", "trainB")):
        ms = []
        for i, p in enumerate(lists[split]):
            mk = np.zeros((72, 80), np.uint8)
            mk[10 + 5 * i:40, 20:60] = 255
            mp = os.path.join(tmp, f"mask_{dom}{i}.png")
            Image.fromarray(mk).save(mp)
            ms.append(mp)
        masks[dom] = ms
    syn_a = lists["trainA"]
    syn_b, syn_m = [], []
    for i, p in enumerate(syn_a):
        arr = np.array(Image.open(p)).copy()
        arr[30:60, 10:50] = 255 - arr[30:60, 10:50]
        pb = os.path.join(tmp, f"syn_b{i}.png")
        Image.fromarray(arr).save(pb)
        syn_b.append(pb)
        mk = np.zeros(arr.shape[:2], np.uint8)
        mk[30:60, 10:50] = 255
        pm = os.path.join(tmp, f"syn_m{i}.png")
        Image.fromarray(mk).save(pm)
        syn_m.append(pm)
    over = dict(recon_mask=1, recon_synth_w=5, synthetic_frequency=1, ratio_disc_gen=1, max_iter=2,
                data_folder_train_a="", data_folder_train_b="",
                data_list_train_a=write_list("la.txt", lists["trainA"]), data_list_train_b=write_list("lb.txt", lists["trainB"]),
                data_list_train_a_seg=write_list("lma.txt", masks["a"]), data_list_train_b_seg=write_list("lmb.txt", masks["b"]),
                data_list_train_a_synth=write_list("sa.txt", syn_a), data_list_train_b_synth=write_list("sb.txt", syn_b),
                data_list_train_b_seg_synth=write_list("sm.txt", syn_m))
    main = _train_main()
    cfg_path = _config(tmp, **over)
    assert main(["--config", cfg_path, "--output_path", tmp]) == 2
    assert os.path.isfile(os.path.join(tmp, "outputs", "tiny", "checkpoints", "gen_00000002.pt"))


def test_train_script_adaptation_heads(tmp_path):
    """config_256's adaptation terms (adv_lambda 6, dfeat_lambda 1) through the training loop at 256x256: the
    classifier-fooling loss in gen_update and a domain_classifier_sr_update every `classif_frequency` iterations
    (train.py:193-207)."""
    tmp = str(tmp_path)
    for k, split in enumerate(("trainA", "trainB", "testA", "testB")):
        _images(os.path.join(tmp, "data", split), 2, 20 + k, size=(272, 264))
    cfg = O.config_256_core()
    cfg["adaptation"].update(adv_lambda=6, dfeat_lambda=1, classif_frequency=1)
    main = _train_main()
    cfg_path = _config(tmp, crop_image_height=256, crop_image_width=256, new_size=256, adaptation=cfg["adaptation"],
                       ratio_disc_gen=1, max_iter=2, image_save_iter=100, snapshot_save_iter=100)
    import munit_b200.trainer as T

    seen = []
    orig = T.MUNIT_Trainer.domain_classifier_sr_update

    def spy(self, *a, **k):
        out = orig(self, *a, **k)
        seen.append((float(self.loss_classifier_sr_update), float(self.loss_classifier_sr)))
        return out

    T.MUNIT_Trainer.domain_classifier_sr_update = spy
    try:
        assert main(["--config", cfg_path, "--output_path", tmp]) == 2
    finally:
        T.MUNIT_Trainer.domain_classifier_sr_update = orig
    assert len(seen) == 2 and all(np.isfinite(v) for pair in seen for v in pair), seen
