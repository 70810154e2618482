"""Hot-path helpers of cc-ai/MUNIT `scripts/utils.py`: get_config (:743-758), get_scheduler (:1066-1090),
weights_init (:1093-1115), get_model_list (:887-908); conv3x3 / conv1x1 / BasicBlock / domainClassifier
(:1238-1392) are re-exported from munit_b200.heads.  FID and segmentation are out of scope."""
from __future__ import annotations

import math
import os

import torch
import yaml
from torch.nn import init
from torch.optim import lr_scheduler


def get_config(config):
    """Parse a yaml config file; defaults `optimizer` to "adam" (utils.py:743-758)."""
    with open(config, "r") as stream:
        conf = yaml.safe_load(stream)
    if "optimizer" not in conf:
        conf["optimizer"] = "adam"
    return conf


def core_config(conf: dict) -> dict:
    """The benchmark configuration "config_256-core" (SURVEY.md D3/D5): the shipped yaml with the terms
    that need unavailable checkpoints / out-of-scope heads switched off."""
    conf = dict(conf)
    conf["semantic_w"] = 0
    conf["recon_mask"] = 0
    conf["domain_adv_w"] = 0
    conf.setdefault("recon_synth_w", 0)
    conf.setdefault("optimizer", "adam")
    ad = dict(conf.get("adaptation") or {})
    for k in ("full_adaptation", "output_classifier_lambda", "output_adv_lambda", "adv_lambda", "dfeat_lambda",
              "sem_seg_lambda"):
        ad[k] = 0
    ad.setdefault("output_classif_freq", 1)
    ad.setdefault("classif_frequency", 15)
    conf["adaptation"] = ad
    return conf


def get_scheduler(optimizer, hyperparameters, iterations=-1):
    """StepLR(step_size, gamma, last_epoch=iterations) or None for a constant policy (utils.py:1066-1090)."""
    if "lr_policy" not in hyperparameters or hyperparameters["lr_policy"] == "constant":
        scheduler = None
    elif hyperparameters["lr_policy"] == "step":
        if iterations != -1:
            for g in optimizer.param_groups:
                g.setdefault("initial_lr", g["lr"])
        scheduler = lr_scheduler.StepLR(optimizer, step_size=hyperparameters["step_size"],
                                        gamma=hyperparameters["gamma"], last_epoch=iterations)
    else:
        return NotImplementedError("learning rate policy [%s] is not implemented", hyperparameters["lr_policy"])
    return scheduler


def weights_init(init_type="gaussian"):
    """utils.py:1093-1115.  The random draw is made on a standard-layout temporary and copied into the
    (channels_last) parameter so a seeded construction yields the reference's values index for index."""

    def init_fun(m):
        classname = m.__class__.__name__
        if (classname.find("Conv") == 0 or classname.find("Linear") == 0) and hasattr(m, "weight"):
            if not isinstance(m.weight, torch.Tensor):
                return
            tmp = torch.empty(m.weight.shape, dtype=m.weight.dtype, device=m.weight.device)
            if init_type == "gaussian":
                init.normal_(tmp, 0.0, 0.02)
            elif init_type == "xavier":
                init.xavier_normal_(tmp, gain=math.sqrt(2))
            elif init_type == "kaiming":
                init.kaiming_normal_(tmp, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(tmp, gain=math.sqrt(2))
            elif init_type == "default":
                tmp.copy_(m.weight.data)
            else:
                assert 0, "Unsupported initialization: {}".format(init_type)
            m.weight.data.copy_(tmp)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)

    return init_fun


def get_model_list(dirname, key):
    """Lexicographically last `*key*.pt` in dirname (utils.py:887-908)."""
    if os.path.exists(dirname) is False:
        return None
    gen_models = [os.path.join(dirname, f) for f in os.listdir(dirname)
                  if os.path.isfile(os.path.join(dirname, f)) and key in f and ".pt" in f]
    if not gen_models:
        return None
    gen_models.sort()
    return gen_models[-1]


# ------------------------------------------------------------------ train.py helpers (utils.py:768-835,1118-1128)
def prepare_sub_folder(output_directory):
    """images/ and checkpoints/ under the run directory; returns (checkpoint_directory, image_directory)."""
    image_directory = os.path.join(output_directory, "images")
    checkpoint_directory = os.path.join(output_directory, "checkpoints")
    for d in (image_directory, checkpoint_directory):
        if not os.path.exists(d):
            print("Creating directory: {}".format(d))
            os.makedirs(d)
    return checkpoint_directory, image_directory


def _write_images(image_outputs, display_image_num, file_name):
    import torchvision.utils as vutils

    rows = [im.expand(-1, 3, -1, -1)[:display_image_num] for im in image_outputs]  # grey -> 3 channels
    grid = vutils.make_grid(torch.cat(rows, 0).data, nrow=display_image_num, padding=0, normalize=True)
    vutils.save_image(grid, file_name, nrow=1)


def write_2images(image_outputs, display_image_num, image_directory, postfix, comet_exp=None):
    """First half of `image_outputs` is the a->b strip, second half b->a (trainer.sample order)."""
    n = len(image_outputs)
    names = ["%s/gen_a2b_%s.jpg" % (image_directory, postfix), "%s/gen_b2a_%s.jpg" % (image_directory, postfix)]
    _write_images(image_outputs[0:n // 2], display_image_num, names[0])
    _write_images(image_outputs[n // 2:n], display_image_num, names[1])
    if comet_exp is not None:
        for nm in names:
            comet_exp.log_image(nm)


class Timer:
    def __init__(self, msg):
        self.msg, self.start_time = msg, None

    def __enter__(self):
        import time

        self.start_time = time.time()

    def __exit__(self, exc_type, exc_value, exc_tb):
        import time

        print(self.msg % (time.time() - self.start_time))


from .heads import BasicBlock, conv1x1, conv3x3, domainClassifier  # noqa: E402,F401  (utils.py:1238-1392)
