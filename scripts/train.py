"""Drop-in for cc-ai/MUNIT scripts/train.py (train.py:36-330) on the B200 kernels: same flags, same config keys,
same outputs (images/, checkpoints/, config.yaml under <output_path>/outputs/<config name>).

Differences, all on the side of doing less: comet_ml is optional (absent -> no experiment logging); the
semantic-segmentation, VGG, FID, domain_adv_w and output-classifier branches are outside the hot path
and exit with a message if the config enables them (the content-feature classifiers of adaptation.adv_lambda /
dfeat_lambda are built: train.py:193-207,261-274); the training loop also runs when semantic_w == 0 (the reference only
loops under semantic_w != 0, train.py:159).  Masks (recon_mask: 1) and synthetic pairs (synthetic_frequency > 0)
are supported through the list-file keys the reference uses."""
from __future__ import print_function

import argparse
import os
import shutil
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from munit_b200 import data as D  # noqa: E402
from munit_b200.trainer import MUNIT_Trainer  # noqa: E402
from munit_b200.utils import Timer, get_config, prepare_sub_folder, write_2images  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--config", type=str, default="configs/config256.yaml", help="Path to the config file.")
parser.add_argument("--output_path", type=str, default=".", help="outputs path")
parser.add_argument("--resume", action="store_true")
parser.add_argument("--trainer", type=str, default="MUNIT", help="MUNIT|UNIT")
parser.add_argument("--project", type=str, default="testing-munit", help="Comet's project_name")
parser.add_argument("--workspace", type=str, default="sunandr", help="Comet's workspace")


def _unsupported(config):
    off = []
    if config.get("semantic_w", 0) != 0:
        off.append("semantic_w (segmentation consistency)")
    if config.get("vgg_w", 0) != 0:
        off.append("vgg_w (perceptual loss)")
    if config.get("domain_adv_w", 0) != 0:
        off.append("domain_adv_w (domain classifier)")
    if config.get("eval_fid", 0) > 0:
        off.append("eval_fid")
    ad = config.get("adaptation", {})
    if any(ad.get(k, 0) for k in ("output_adv_lambda", "output_classifier_lambda", "sem_seg_lambda")):
        off.append("adaptation.output_* / sem_seg_lambda heads")
    return off


def _mask_loader(config, dom):
    """Per-domain masks for recon_mask: 1 -- a list file of mask images aligned index-by-index with the train
    list of the domain (data_list_train_<dom>_seg, the key the reference's mask loader reads, train.py:83-106)."""
    key = f"data_list_train_{dom}_seg"
    if key not in config:
        sys.exit(f"recon_mask: 1 needs {key}")
    return D.read_list(config[key])


def main(argv=None):
    opts = parser.parse_args(argv)
    if opts.trainer != "MUNIT":
        sys.exit("Only support MUNIT")
    try:
        from comet_ml import Experiment  # optional

        comet_exp = Experiment(workspace=opts.workspace, project_name=opts.project)
    except Exception:
        comet_exp = None
    config = get_config(opts.config)
    off = _unsupported(config)
    if off:
        sys.exit("outside the B200 hot path (set to 0 or use munit_b200.utils.core_config): " + ", ".join(off))
    max_iter, display_size = config["max_iter"], config["display_size"]
    config["vgg_model_path"] = opts.output_path
    config.setdefault("recon_synth_w", 0)
    config.setdefault("recon_mask", 0)
    trainer = MUNIT_Trainer(config)
    trainer.cuda()
    # opt-in extension key (not in the reference's configs): gen_update reuses the generator pass of the dis_update
    # that preceded it on the same batch (exactly transparent; pays off when ratio_disc_gen == 1 and guided == 1)
    trainer.reuse_forward = bool(config.get("reuse_forward", 0))
    train_loader_a, train_loader_b, test_loader_a, test_loader_b = D.get_all_data_loaders(config)
    synthetic_loader = None
    if config.get("synthetic_frequency", 0) > 0:
        synthetic_loader = D.get_synthetic_data_loader(
            config["data_list_train_a_synth"], config["data_list_train_b_synth"], config["data_list_train_b_seg_synth"],
            config["batch_size"], True, new_size=config["new_size"], height=config["crop_image_height"],
            width=config["crop_image_width"], num_workers=config["num_workers"])
    use_masks = config["recon_mask"] == 1
    if use_masks:
        # image + mask with one shared crop / flip: the paired dataset with the image as both "a" and "b"
        mk = lambda dom: D._loader(D.PairedWithMask(  # noqa: E731
            [os.path.join(config.get(f"data_folder_train_{dom}", ""), p) for p in D.read_list(config[f"data_list_train_{dom}"])],
            [os.path.join(config.get(f"data_folder_train_{dom}", ""), p) for p in D.read_list(config[f"data_list_train_{dom}"])],
            _mask_loader(config, dom), config["new_size"], config["crop_image_height"], config["crop_image_width"], True),
            config["batch_size"], True, config["num_workers"])
        train_loader_a, train_loader_b = mk("a"), mk("b")

    def first_images(loader):
        items = [loader.dataset[i] for i in range(display_size)]
        return torch.stack([it[0] if isinstance(it, (tuple, list)) else it for it in items]).cuda()

    train_display_images_a, train_display_images_b = first_images(train_loader_a), first_images(train_loader_b)
    test_display_images_a, test_display_images_b = first_images(test_loader_a), first_images(test_loader_b)

    model_name = os.path.splitext(os.path.basename(opts.config))[0]
    output_directory = os.path.join(opts.output_path + "/outputs", model_name)
    checkpoint_directory, image_directory = prepare_sub_folder(output_directory)
    shutil.copy(opts.config, os.path.join(output_directory, "config.yaml"))
    iterations = trainer.resume(checkpoint_directory, hyperparameters=config) if opts.resume else 0
    trainer.iterations = iterations
    ratio = config.get("ratio_disc_gen", 1)
    synth_iter = iter(synthetic_loader) if synthetic_loader is not None else None
    while True:
        for batch_a, batch_b in zip(train_loader_a, train_loader_b):
            with Timer("Elapsed time in update s: %f"):
                trainer.update_learning_rate()
                if use_masks:
                    images_a, mask_a = batch_a[0].cuda(non_blocking=True), batch_a[2].cuda(non_blocking=True)
                    images_b, mask_b = batch_b[0].cuda(non_blocking=True), batch_b[2].cuda(non_blocking=True)
                else:
                    images_a, images_b = batch_a.cuda(non_blocking=True), batch_b.cuda(non_blocking=True)
                    mask_a = mask_b = None
                trainer.dis_update(images_a, images_b, config, comet_exp)
                if (iterations + 1) % ratio == 0:
                    trainer.gen_update(images_a, images_b, config, mask_a, mask_b, comet_exp)
                # content-feature classifier update on real pairs (train.py:193-207)
                classif_now = (trainer.use_classifier_sr
                               and (iterations + 1) % config["adaptation"]["classif_frequency"] == 0)
                if classif_now:
                    trainer.domain_classifier_sr_update(images_a, images_b, False, config["adaptation"]["dfeat_lambda"],
                                                        iterations + 1, comet_exp)
                if synth_iter is not None and iterations % config["synthetic_frequency"] == 0:
                    try:
                        images_as, images_bs, mask_s = next(synth_iter)
                    except StopIteration:
                        synth_iter = iter(synthetic_loader)
                        images_as, images_bs, mask_s = next(synth_iter)
                    images_as, images_bs, mask_s = images_as.cuda(), images_bs.cuda(), mask_s.cuda()
                    trainer.dis_update(images_as, images_bs, config, comet_exp)
                    trainer.gen_update(images_as, images_bs, config, mask_s, mask_s, comet_exp, True)
                    if classif_now:  # ... and on the synthetic pair (train.py:261-274)
                        trainer.domain_classifier_sr_update(images_as, images_bs, True,
                                                            config["adaptation"]["dfeat_lambda"], iterations + 1, comet_exp)
                torch.cuda.synchronize()
            if (iterations + 1) % config["image_save_iter"] == 0:
                with torch.no_grad():
                    test_out = trainer.sample(test_display_images_a, test_display_images_b)
                    train_out = trainer.sample(train_display_images_a, train_display_images_b)
                write_2images(test_out, display_size, image_directory, "test_%08d" % (iterations + 1), comet_exp)
                write_2images(train_out, display_size, image_directory, "train_%08d" % (iterations + 1), comet_exp)
            if (iterations + 1) % config["image_display_iter"] == 0:
                with torch.no_grad():
                    out = trainer.sample(train_display_images_a, train_display_images_b)
                write_2images(out, display_size, image_directory, "train_current", comet_exp)
            if (iterations + 1) % config["snapshot_save_iter"] == 0:
                trainer.save(checkpoint_directory, iterations)
            iterations += 1
            trainer.iterations = iterations
            if iterations >= max_iter:
                print("Finish training")
                return iterations


if __name__ == "__main__":
    main()
