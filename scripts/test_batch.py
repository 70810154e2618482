"""Drop-in for the MUNIT branch of cc-ai/MUNIT scripts/test_batch.py (test_batch.py:89-208): for every image of
--input_folder encode once and decode --num_style random style codes (gen_state 0 checkpoints {"a","b"}).
The reference loops decode() per style at batch 1; every norm is per-sample, so the styles are batched here."""
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
parser.add_argument("--config", type=str, help="net configuration")
parser.add_argument("--input_folder", type=str, help="input image folder")
parser.add_argument("--output_folder", type=str, help="output image folder")
parser.add_argument("--checkpoint", type=str, help="checkpoint of autoencoders")
parser.add_argument("--a2b", type=int, help="1 for a2b and others for b2a", default=1)
parser.add_argument("--seed", type=int, default=1, help="random seed")
parser.add_argument("--num_style", type=int, default=10, help="number of styles to sample")
parser.add_argument("--synchronized", action="store_true", help="whether use synchronized style code or not")
parser.add_argument("--output_only", action="store_true", help="whether only save the output images or also save the input images")
parser.add_argument("--output_path", type=str, default=".", help="path for logs, checkpoints, and VGG model weight")
parser.add_argument("--trainer", type=str, default="MUNIT", help="MUNIT")


def translate(encode, decode, images, styles):
    """images [B,3,H,W], styles [S,style_dim,1,1] -> [S][B,3,H,W]: one content encode, S decodes (batched over B).
    `encode` returns the content code only (the reference's gen.encode also runs the style encoder on the content
    image and discards the result, test_batch.py:154)."""
    content = encode(images)
    return [decode(content, styles[j:j + 1].expand(images.shape[0], -1, -1, -1).contiguous())
            for j in range(styles.shape[0])]


def main():
    import torchvision.utils as vutils
    from PIL import Image
    from torchvision import transforms

    opts = parser.parse_args()
    if opts.trainer != "MUNIT":
        sys.exit("Only support MUNIT")
    torch.manual_seed(opts.seed)
    torch.cuda.manual_seed(opts.seed)
    config = core_config(get_config(opts.config))
    config["gen_state"] = 0
    style_dim = config["gen"]["style_dim"]
    trainer = MUNIT_Trainer(config)
    state_dict = torch.load(opts.checkpoint, map_location="cpu")
    trainer.gen_a.load_state_dict(state_dict["a"])
    trainer.gen_b.load_state_dict(state_dict["b"])
    trainer.cuda()
    trainer.eval()
    enc = trainer.gen_a.enc_content if opts.a2b else trainer.gen_b.enc_content
    encode = lambda x: enc.forward_act(x, 1)
    decode = trainer.gen_b.decode if opts.a2b else trainer.gen_a.decode
    tf = transforms.Compose([transforms.Resize(config.get("new_size_a", config["new_size"])), transforms.ToTensor(),
                             transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    style_fixed = torch.randn(opts.num_style, style_dim, 1, 1).cuda()
    # the reference iterates a torch DataLoader here (test_batch.py:110-112,148): creating its iterator draws one int64
    # base seed from the host generator before the per-image style codes are drawn -- same draw, same codes
    torch.empty((), dtype=torch.int64).random_()
    with torch.no_grad():
        for i, path in enumerate(sorted(glob.glob(os.path.join(opts.input_folder, "*")))):
            x = tf(Image.open(path).convert("RGB")).unsqueeze(0)
            # cut to a multiple of 4 (2 stride-2 layers): the reference's output has 4 * floor(W / 4) columns too
            h, w = x.shape[2] // 4 * 4, x.shape[3] // 4 * 4
            images = x[:, :, :h, :w].contiguous().cuda()
            style = style_fixed if opts.synchronized else torch.randn(opts.num_style, style_dim, 1, 1).cuda()
            for j, out in enumerate(translate(encode, decode, images, style)):
                p = os.path.join(opts.output_folder + "_%02d" % j, os.path.basename(path))
                os.makedirs(os.path.dirname(p), exist_ok=True)
                vutils.save_image(((out + 1) / 2.0).data, p, padding=0, normalize=True)
            if not opts.output_only:
                os.makedirs(opts.output_folder, exist_ok=True)
                vutils.save_image(images.data, os.path.join(opts.output_folder, "input{:03d}.jpg".format(i)), padding=0,
                                  normalize=True)


if __name__ == "__main__":
    main()
