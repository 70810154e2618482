"""Input-pipeline throughput: images/s of the reference's host pipeline (decode -> flip -> resize ->
crop -> ToTensor -> Normalize on the workers, fp32 batches over PCIe) against the GPU-tail pipeline (workers stop at
the resized uint8 image; crop + ToTensor + Normalize in munit_u8_crop_normalize), on synthetic JPEGs.

    python tools/bench_loader.py [n_images] [workers] [out.json]

The 8-GPU training step consumes 2 * 8 * 8 images per 38 ms = ~3400 images/s; this says how many host cores that needs."""
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from munit_b200 import data as D  # noqa: E402


def main():
    from PIL import Image

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 4)
    d = tempfile.mkdtemp()
    g = torch.Generator().manual_seed(0)
    base = torch.rand(96, 128, 3, generator=g)
    for i in range(n):  # smooth random images (JPEG-compressible), 1024 x 768 like web photos
        img = torch.nn.functional.interpolate(base.roll(i, 1).permute(2, 0, 1)[None], size=(768, 1024), mode="bilinear")[0]
        Image.fromarray((img.permute(1, 2, 0) * 255).byte().numpy()).save(os.path.join(d, f"im{i:04d}.jpg"), quality=90)
    kw = dict(batch_size=8, train=True, new_size=256, height=256, width=256, num_workers=workers, crop=True)
    res = dict(images=n, workers=workers, cores=os.cpu_count(), source="1024x768 JPEG q90 -> Resize(256) -> 256x256 crop")
    for name, mk in (("host_fp32", lambda: D.get_data_loader_folder(d, **kw)),
                     ("gpu_tail_u8", lambda: D.get_gpu_data_loader_folder(d, **kw))):
        ld = mk()
        for epoch in range(2):  # epoch 0 warms the workers / page cache
            t0 = time.time()
            cnt = 0
            for b in ld:
                if not b.is_cuda:
                    b = b.cuda(non_blocking=True)
                cnt += b.shape[0]
            torch.cuda.synchronize()
            dt = time.time() - t0
        res[name] = dict(images_per_s=cnt / dt, seconds=dt)
        print(name, "%.0f images/s" % (cnt / dt), flush=True)
    need = 2 * 8 * 8 / 0.038
    res["needed_for_8_gpus_images_per_s"] = need
    res["note"] = ("both pipelines are bound by the PIL JPEG decode + resize on the host workers; the GPU tail removes the float "
                   "conversion and 3/4 of the PCIe bytes")
    print(json.dumps(res))
    if len(sys.argv) > 3:
        json.dump(res, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
