"""Drop-in for cc-ai/MUNIT scripts/test.py (test.py:20-129): translate every image of --input with the style of
--style through the shared-style generator (gen_state 1), on the B200 kernels.  Same CLI flags."""
from __future__ import print_function

import argparse
import glob
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from munit_b200.trainer import MUNIT_Trainer  # noqa: E402
from munit_b200.utils import core_config, get_config  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--config", type=str, help="network configuration file")
parser.add_argument("--input", type=str, help="directory of input images")
parser.add_argument("--output_folder", type=str, help="output image directory")
parser.add_argument("--checkpoint", type=str, help="checkpoint of generator")
parser.add_argument("--style", type=str, default="", help="style image path")
parser.add_argument("--seed", type=int, default=10, help="random seed")
parser.add_argument("--synchronized", action="store_true", help="whether use synchronized style code or not")
parser.add_argument("--save_input", action="store_true", help="also save the (denormalised) inputs")
parser.add_argument("--output_path", type=str, default=".", help="path for logs, checkpoints, and VGG model weight")


def main():
    import torchvision.utils as vutils
    from PIL import Image
    from torchvision import transforms

    opts = parser.parse_args()
    torch.manual_seed(opts.seed)
    torch.cuda.manual_seed(opts.seed)
    os.makedirs(opts.output_folder, exist_ok=True)
    config = core_config(get_config(opts.config))  # inference needs none of the auxiliary heads
    trainer = MUNIT_Trainer(config)
    try:
        state_dict = torch.load(opts.checkpoint, map_location="cpu")
        trainer.gen.load_state_dict(state_dict["2"])
    except Exception:
        sys.exit("Cannot load the checkpoints")
    trainer.cuda()
    trainer.eval()
    new_size = config["new_size"]
    list_non_flooded = sorted(glob.glob(opts.input + "*"))
    if len(list_non_flooded) == 0:
        sys.exit("Image list is empty. Please ensure opts.input ends with a /")
    with torch.no_grad():
        transform = transforms.Compose([transforms.Resize(new_size), transforms.ToTensor(),
                                        transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])

        def load(path, multiple):
            """Resize(new_size) -> ToTensor -> Normalize (test.py:88-94), then the image is cut (right / bottom) to the
            largest extent the strided encoders map exactly: a multiple of 4 for the content path (2 stride-2 layers)
            and of 16 for the style encoder (4).  The reference feeds the resized image as is and its stride-2
            convolutions silently ignore the odd last row / column, its up-sampling decoder returns 4 * floor(W / 4)
            columns (341 -> 340): output sizes are therefore the same, and only the few output columns whose receptive
            field reaches the dropped input column differ (tests/test_inference_gpu.py compares against the
            reference's outputs on its own demo images)."""
            x = transform(Image.open(path).convert("RGB")).unsqueeze(0)
            h, w = x.shape[2] // multiple * multiple, x.shape[3] // multiple * multiple
            if h == 0 or w == 0:
                sys.exit("image %s is smaller than %d pixels after Resize(%d)" % (path, multiple, new_size))
            return x[:, :, :h, :w].contiguous().cuda()

        # (the reference calls gen.encode(), which also runs the style encoder on the content image and the content
        # encoder on the style image and discards both, test.py:100,114; only the halves that are used run here)
        s_b = trainer.gen.enc_style(load(opts.style, 16))
        for j, path_xa in enumerate(list_non_flooded):
            x_a = load(path_xa, 4)
            if opts.save_input:
                vutils.save_image(((x_a + 1) / 2.0).data, os.path.join(opts.output_folder, "input{:03d}.jpg".format(j)),
                                  padding=0, normalize=True)
            c_a = trainer.gen.enc1_content.forward_act(x_a, 1)
            x_ab = trainer.gen.decode(c_a, s_b, 2)
            outputs = (x_ab + 1) / 2.0
            vutils.save_image(outputs.data, os.path.join(opts.output_folder, "output{:03d}.jpg".format(j)), padding=0,
                              normalize=True)


if __name__ == "__main__":
    main()
