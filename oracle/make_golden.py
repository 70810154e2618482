"""Generate tests/golden/*.pt from the reference itself -- run in the build container:

    python -m oracle.make_golden

Weights are produced by munit_oracle.init_state_dict(seed) (so they can be rebuilt on the
GPU box without shipping 100+ MB) and loaded into the reference modules; inputs are seeded;
the fixtures hold the reference's outputs / gradients / losses / post-step weights
(small tensors whole, big ones as strided samples + sums).
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import munit_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def summarize(t: torch.Tensor, n=64):
    """Compact fingerprint of a big tensor: sum, abs-sum, and a strided sample."""
    f = t.detach().reshape(-1).double()
    step = max(1, f.numel() // n)
    return dict(shape=tuple(t.shape), sum=float(f.sum()), asum=float(f.abs().sum()),
                sample=f[::step][:n].float().clone(), step=step)


def seeded_images(seed, b, h, w):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, h, w, generator=g) * 2 - 1, torch.rand(b, 3, h, w, generator=g) * 2 - 1


def layers_fixture():
    """Per-layer forward/backward of the reference modules (networks.py) on tiny inputs."""
    networks, _ = ref_loader.load()
    torch.manual_seed(7)
    fx = {}
    # Conv2dBlock variants: (cin,cout,k,s,p,norm,act)
    for name, (cin, cout, k, s, p, norm, act, hw) in {
        "c7_in_relu": (3, 16, 7, 1, 3, "in", "relu", 16),
        "c4s2_none_lrelu": (16, 24, 4, 2, 1, "none", "lrelu", 16),
        "c3_in_none": (16, 16, 3, 1, 1, "in", "none", 12),
        "c5_ln_relu": (24, 16, 5, 1, 2, "ln", "relu", 12),
        "c7_none_tanh": (16, 3, 7, 1, 3, "none", "tanh", 16),
    }.items():
        m = networks.Conv2dBlock(cin, cout, k, s, p, norm=norm, activation=act, pad_type="reflect")
        x = torch.randn(2, cin, hw, hw, requires_grad=True)
        y = m(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        fx[name] = dict(args=(cin, cout, k, s, p, norm, act), sd={k_: v.detach().clone() for k_, v in m.state_dict().items()},
                        x=x.detach().clone(), y=y.detach().clone(), gy=gy, gx=x.grad.clone(),
                        gw=m.conv.weight.grad.clone(), gb=m.conv.bias.grad.clone(),
                        gnorm={n: p_.grad.clone() for n, p_ in m.named_parameters() if n.startswith("norm")})
    # AdaIN
    m = networks.AdaptiveInstanceNorm2d(64)
    x = torch.randn(3, 64, 8, 8, requires_grad=True)
    w = torch.randn(3 * 64, requires_grad=True)
    b = torch.randn(3 * 64, requires_grad=True)
    m.weight, m.bias = w, b
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    fx["adain"] = dict(x=x.detach().clone(), w=w.detach().clone(), b=b.detach().clone(), y=y.detach().clone(),
                       gy=gy, gx=x.grad.clone(), gw=w.grad.clone(), gb=b.grad.clone())
    # LayerNorm
    m = networks.LayerNorm(32)
    x = torch.randn(2, 32, 8, 8, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    fx["ln"] = dict(x=x.detach().clone(), gamma=m.gamma.detach().clone(), beta=m.beta.detach().clone(),
                    y=y.detach().clone(), gy=gy, gx=x.grad.clone(), ggamma=m.gamma.grad.clone(), gbeta=m.beta.grad.clone())
    # avg-pool pyramid divisor
    d = networks.MsImageDis(3, O.config_256_core()["dis"])
    x = torch.randn(1, 3, 16, 16)
    fx["avgpool"] = dict(x=x, y=d.downsample(x))
    torch.save(fx, os.path.join(OUT, "layers.pt"))
    print("layers.pt", {k: None for k in fx})


def nets_fixture():
    """Whole-network forwards on seeded weights: AdaINGen_double encode/decode, MsImageDis."""
    networks, _ = ref_loader.load()
    cfg = O.config_256_core()
    gsd = O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 11, "kaiming")
    dsd = O.init_state_dict(O.dis_spec(cfg["dis"], 3), 12, "gaussian")
    gen = networks.AdaINGen_double(3, cfg["gen"])
    gen.load_state_dict(gsd)
    dis = networks.MsImageDis(3, cfg["dis"])
    dis.load_state_dict(dsd)
    x_a, x_b = seeded_images(1234, 2, 64, 64)
    with torch.no_grad():
        c_a, s_a = gen.encode(x_a, 1)
        c_b, s_b = gen.encode(x_b, 2)
        x_ab = gen.decode(c_a, s_b, 2)
        x_ba = gen.decode(c_b, s_a, 1)
        ap = gen.get_adain_param(s_b, 2)
        d_out = dis(x_ab)
        dl = dis.calc_dis_loss(x_ab, x_b)
        gl = dis.calc_gen_loss(x_ab)
    # single-generator AdaINGen (gen_state 0) too
    gsd0 = O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 13, "kaiming")
    g0 = networks.AdaINGen(3, cfg["gen"])
    g0.load_state_dict(gsd0)
    with torch.no_grad():
        c0, s0 = g0.encode(x_a)
        torch.manual_seed(5)
        s_rand = torch.randn(2, cfg["gen"]["style_dim"], 1, 1)
        y0 = g0.decode(c0, s_rand)
    fx = dict(seeds=dict(gen=11, dis=12, gen0=13, img=1234), hw=64, b=2,
              c_a=summarize(c_a, 256), s_a=s_a.clone(), s_b=s_b.clone(), x_ab=x_ab.clone(), x_ba=x_ba.clone(),
              adain_params=summarize(ap, 256), d_out=[o.clone() for o in d_out], dis_loss=float(dl), gen_loss=float(gl),
              s_rand=s_rand, y0=y0.clone(), c0=summarize(c0, 256), s0=s0.clone())
    torch.save(fx, os.path.join(OUT, "nets.pt"))
    print("nets.pt ok")


def classifier_fixture():
    """domainClassifier (utils.py:1370-1392) forward / backward in train mode (BatchNorm batch statistics + running
    update), then an eval-mode forward on the updated running statistics."""
    ref_loader.load()
    import utils as ref_utils

    sd = O.init_classifier_state_dict(31)
    m = ref_utils.domainClassifier(256)
    m.load_state_dict(sd)
    m.train()
    g = torch.Generator().manual_seed(32)
    x = (torch.randn(2, 256, 64, 64, generator=g) * 1.5).requires_grad_(True)
    y = m(x)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    after = {k: v.detach().clone() for k, v in m.state_dict().items() if "running" in k or "tracked" in k}
    grads = {n: p_.grad.clone() if p_.numel() <= 4096 else summarize(p_.grad, 256) for n, p_ in m.named_parameters()}
    m.eval()
    with torch.no_grad():
        y_eval = m(x.detach())
    fx = dict(seed=31, x_seed=32, y=y.detach().clone(), gy=gy, gx=summarize(x.grad, 1024), gx_absmax=float(x.grad.abs().max()),
              grads=grads, running_after=after, y_eval=y_eval.clone())
    torch.save(fx, os.path.join(OUT, "classifier.pt"))
    print("classifier.pt ok", y.detach().reshape(-1), y_eval.reshape(-1))


def step_fixture(tag, optimizer, gen_state, guided, hw, b, n_steps, masked_synth=False, adaptation=False, style_dim=None):
    """dis_update + gen_update on the shimmed reference trainer (trainer.py:336-561,1133-1186).  masked_synth:
    recon_mask = 1 with masks and synth=True with recon_synth_w > 0 (trainer.py:452-488).  adaptation: the
    config_256 values adaptation.adv_lambda = 6, dfeat_lambda = 1 (content-feature classifiers, trainer.py:162-179,
    521-525) and one domain_classifier_sr_update (trainer.py:1237-1265) after gen_update, as train.py:193-207."""
    cfg = O.config_256_core(optimizer=optimizer, gen_state=gen_state, guided=guided,
                            crop_image_height=hw, crop_image_width=hw)
    if adaptation:
        cfg["adaptation"].update(adv_lambda=6, dfeat_lambda=1)
    if style_dim is not None:
        cfg["gen"]["style_dim"] = style_dim
    if masked_synth:
        cfg["recon_mask"], cfg["recon_synth_w"] = 1, 5
        torch.cuda.FloatTensor = torch.FloatTensor  # shim 3: trainer.py:456 casts the alignment mask to a CUDA type
    torch.manual_seed(0)
    t = ref_loader.make_trainer(cfg)
    if gen_state == 1:
        gsd = O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 21, "kaiming")
        t.gen.load_state_dict(gsd)
    else:
        ga = O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 21, "kaiming")
        gb = O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 22, "kaiming")
        t.gen_a.load_state_dict(ga)
        t.gen_b.load_state_dict(gb)
    t.dis_a.load_state_dict(O.init_state_dict(O.dis_spec(cfg["dis"], 3), 23, "gaussian"))
    t.dis_b.load_state_dict(O.init_state_dict(O.dis_spec(cfg["dis"], 3), 24, "gaussian"))
    if adaptation:
        t.domain_classifier_sr_a.load_state_dict(O.init_classifier_state_dict(25))
        t.domain_classifier_sr_b.load_state_dict(O.init_classifier_state_dict(26))
    x_a, x_b = seeded_images(1234, b, hw, hw)
    mask_a = mask_b = None
    if masked_synth:
        x_a, x_b, mask_a, mask_b = O.synthetic_pair(1234, b, hw)
    torch.manual_seed(99)  # style-code stream for the updates (trainer.py:366-367,1146-1147)
    steps = []
    for it in range(n_steps):
        t.iterations = it
        t.update_learning_rate()
        extra = {}
        if adaptation:
            # (train.py:193-207 runs this after gen_update; here it goes first so that the classifier update is
            # pinned on the initial generator and not on weights that carry Adam's sign-amplified fp noise)
            t.domain_classifier_sr_update(x_a, x_b, False, cfg["adaptation"]["dfeat_lambda"], it + 1)
            cls = {"a": t.domain_classifier_sr_a, "b": t.domain_classifier_sr_b}
            extra = dict(cls_g={f"{cn}/{n}": summarize(p.grad, 16) for cn, c in cls.items() for n, p in c.named_parameters()},
                         cls_w={f"{cn}/{n}": summarize(p.data, 16) for cn, c in cls.items() for n, p in c.named_parameters()})
        t.dis_update(x_a, x_b, cfg)
        dis_g = {f"a/{n}": summarize(p.grad, 16) for n, p in t.dis_a.named_parameters()}
        if masked_synth:
            t.gen_update(x_a, x_b, cfg, mask_a, mask_b, synth=True)
        else:
            t.gen_update(x_a, x_b, cfg)
        rec = {k: float(getattr(t, k)) for k in dir(t) if k.startswith("loss_") and torch.is_tensor(getattr(t, k))}
        gens = {"": t.gen} if gen_state == 1 else {"a": t.gen_a, "b": t.gen_b}
        gen_g = {f"{gn}/{n}": summarize(p.grad, 16) for gn, g in gens.items() for n, p in g.named_parameters()}
        gen_w = {f"{gn}/{n}": summarize(p.data, 16) for gn, g in gens.items() for n, p in g.named_parameters()}
        dis_w = {f"a/{n}": summarize(p.data, 16) for n, p in t.dis_a.named_parameters()}
        dis_w.update({f"b/{n}": summarize(p.data, 16) for n, p in t.dis_b.named_parameters()})
        if adaptation:
            extra["cls_running"] = {f"{cn}/{n}": v.detach().clone() for cn, c in cls.items()
                                    for n, v in c.state_dict().items() if "running" in n or "tracked" in n}
        steps.append(dict(losses=rec, dis_grads=dis_g, gen_grads=gen_g, gen_w=gen_w, dis_w=dis_w, **extra))
        print(tag, it, {k: round(v, 5) for k, v in rec.items()})
    fx = dict(cfg=cfg, seeds=dict(gen=21, gen_b=22, dis_a=23, dis_b=24, cls_a=25, cls_b=26, img=1234, style=99), hw=hw,
              b=b, steps=steps, masked_synth=masked_synth, adaptation=adaptation)
    torch.save(fx, os.path.join(OUT, f"step_{tag}.pt"))


def _pool(t, k):
    """Low-resolution fingerprint of an image batch (average pooling by k), fp16."""
    import torch.nn.functional as F
    return F.avg_pool2d(t.detach().float(), k).half()


def autocast_grad_cosines():
    """How far does the REFERENCE ITSELF move when only its convolutions become bf16 tensor-core convolutions?  One
    dis_update + gen_update of the unmodified reference trainer with nn.Conv2d patched to round its input, weight and
    output to bf16 (fp32 accumulation, everything else fp32 -- torch.autocast("cpu") cannot run the reference: AdaIN's
    F.batch_norm rejects the mixed dtypes it produces) against the same update in fp32, same seeded weights, inputs and
    style codes: per-tensor gradient cosine / norm ratio and the losses.  tests/test_trainer_gpu.py prints this beside
    the B200 figures: it is the floor any bf16 convolution pipeline has against the fp32 path."""
    import torch.nn.functional as F

    cfg = O.config_256_core(crop_image_height=64, crop_image_width=64)
    out = {}
    orig = torch.nn.Conv2d._conv_forward
    rnd = lambda t, g: O._RoundSTE.apply(t, g)

    def bf16_conv(self, x, w, b):
        return rnd(F.conv2d(rnd(x, True), rnd(w, False), b, self.stride, self.padding, self.dilation, self.groups), True)

    for mode in ("fp32", "bf16"):
        torch.manual_seed(0)
        t = ref_loader.make_trainer(cfg)
        t.gen.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 21, "kaiming"))
        t.dis_a.load_state_dict(O.init_state_dict(O.dis_spec(cfg["dis"], 3), 23, "gaussian"))
        t.dis_b.load_state_dict(O.init_state_dict(O.dis_spec(cfg["dis"], 3), 24, "gaussian"))
        x_a, x_b = seeded_images(1234, 2, 64, 64)
        torch.manual_seed(99)
        t.iterations = 0
        torch.nn.Conv2d._conv_forward = bf16_conv if mode == "bf16" else orig
        try:
            t.dis_update(x_a, x_b, cfg)
            dis_g = {f"dis_a/{n}": p.grad.detach().float().clone() for n, p in t.dis_a.named_parameters()}
            t.gen_update(x_a, x_b, cfg)
        finally:
            torch.nn.Conv2d._conv_forward = orig
        grads = {f"/{n}": p.grad.detach().float().clone() for n, p in t.gen.named_parameters() if p.grad is not None}
        grads.update(dis_g)
        losses = {k: float(getattr(t, k)) for k in dir(t) if k.startswith("loss_") and torch.is_tensor(getattr(t, k))}
        out[mode] = (grads, losses)
    cos, ratio = {}, {}
    for k, g32 in out["fp32"][0].items():
        g16 = out["bf16"][0][k]
        if float(g32.norm()) < 1e-7:
            continue
        cos[k] = float(torch.dot(g16.reshape(-1), g32.reshape(-1)) / (g16.norm() * g32.norm() + 1e-30))
        ratio[k] = float(g16.norm() / g32.norm())
    res = dict(cos=cos, ratio=ratio, losses_fp32=out["fp32"][1], losses_bf16=out["bf16"][1])
    srt = sorted(cos.values())
    print("reference with bf16 convolutions vs fp32: grad cosine worst %.4f median %.4f" % (srt[0], srt[len(srt) // 2]))
    torch.save(res, os.path.join(OUT, "autocast_ref.pt"))


def sample_fixture():
    """MUNIT_Trainer.sample / sample_fid / forward of the reference (trainer.py:307-334,773-928,1087-1131) on seeded
    weights: gen_state 1 / guided 1 and gen_state 0 / guided 0 (random second style: host RNG after manual_seed)."""
    fx = {}
    for tag, gs, gd in (("g1", 1, 1), ("g0", 0, 0)):
        cfg = O.config_256_core(gen_state=gs, guided=gd, crop_image_height=64, crop_image_width=64, display_size=3)
        torch.manual_seed(5)
        t = ref_loader.make_trainer(cfg)
        if gs == 1:
            t.gen.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 21, "kaiming"))
        else:
            t.gen_a.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 21, "kaiming"))
            t.gen_b.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 22, "kaiming"))
        x_a, x_b = seeded_images(77, 3, 64, 64)
        torch.manual_seed(123)
        with torch.no_grad():
            outs = t.sample(x_a, x_b)
            rng_after = torch.get_rng_state()
            fid = t.sample_fid(x_a, x_b) if gd == 1 else None
            fwd = t.forward(x_a, x_b)
        fx[tag] = dict(cfg=cfg, sample=[o.detach().clone() for o in outs], rng_after=rng_after,
                       sample_fid=None if fid is None else fid.detach().clone(), forward=[o.detach().clone() for o in fwd],
                       s_a=t.s_a.clone(), s_b=t.s_b.clone())
        print("sample", tag, [tuple(o.shape) for o in outs])
    torch.save(fx, os.path.join(OUT, "sample.pt"))


def demo_fixture():
    """The reference's shipped demo images (input_folder/, Style_Image/) through the test.py flow (test.py:86-129,
    gen_state 1, style image) and the test_batch.py flow (test_batch.py:146-164, gen_state 0, random styles) of the
    reference generators on seeded weights.  The images are stored already resized to new_size = 256 (what
    transforms.Resize(256) makes of them: the inference scripts' first step, a no-op on the stored files); the
    reference outputs as 4x average-pooled fingerprints."""
    from PIL import Image
    from torchvision import transforms

    networks, _ = ref_loader.load()
    demo = os.path.join(OUT, "demo")
    os.makedirs(os.path.join(demo, "input_folder"), exist_ok=True)
    os.makedirs(os.path.join(demo, "Style_Image"), exist_ok=True)
    src = [("input_folder/demo_image1.jpg", "input_folder/demo_image1.png"),
           ("input_folder/demo_image2.jpg", "input_folder/demo_image2.png"),
           ("input_folder/demo_image3.png", "input_folder/demo_image3.png"),
           ("Style_Image/style_image.png", "Style_Image/style_image.png")]
    for a, b in src:
        im = transforms.Resize(256)(Image.open(os.path.join(ref_loader.REF_ROOT, a)).convert("RGB"))
        im.save(os.path.join(demo, b), optimize=True)
    tf = transforms.Compose([transforms.Resize(256), transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    load = lambda rel: tf(Image.open(os.path.join(demo, rel)).convert("RGB")).unsqueeze(0)
    cfg = O.config_256_core()
    fx = dict(seeds=dict(gen=41, gen_a=42, gen_b=43, style=1), files=[b for _, b in src[:3]], pool=4)
    # ---- test.py: style of the style image, generator "2" checkpoint
    gen = networks.AdaINGen_double(3, cfg["gen"])
    gen.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 41, "kaiming"))
    outs = []
    with torch.no_grad():
        _, s_b = gen.encode(load("Style_Image/style_image.png"), 2)
        for rel in fx["files"]:
            c_a, _ = gen.encode(load(rel), 1)
            outs.append(_pool((gen.decode(c_a, s_b, 2) + 1) / 2.0, 4))
    fx["test_py"] = dict(s_b=s_b.clone(), outputs=outs, shapes=[tuple(o.shape) for o in outs])
    # ---- test_batch.py: 3 random styles per image, a2b
    ga, gb = networks.AdaINGen(3, cfg["gen"]), networks.AdaINGen(3, cfg["gen"])
    ga.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 42, "kaiming"))
    gb.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 43, "kaiming"))
    # the script's own order of host-RNG use (test_batch.py:89-164): seed, trainer construction (weight init + display
    # codes), checkpoint load, style_fixed, the DataLoader iterator's base-seed draw, then one style draw per image
    num_style = 3
    torch.manual_seed(1)
    ref_loader.make_trainer(O.config_256_core(gen_state=0, guided=0))
    style_fixed = torch.randn(num_style, cfg["gen"]["style_dim"], 1, 1)  # unused unless --synchronized
    torch.empty((), dtype=torch.int64).random_()  # what creating the DataLoader iterator draws (torch.utils.data)
    tb = []
    with torch.no_grad():
        for rel in fx["files"]:
            img = load(rel)
            style = torch.randn(num_style, cfg["gen"]["style_dim"], 1, 1)
            content, _ = ga.encode(img)
            tb.append([_pool((gb.decode(content, style[j:j + 1]) + 1) / 2.0, 4) for j in range(num_style)])
    fx["test_batch"] = dict(num_style=num_style, outputs=tb)
    torch.save(fx, os.path.join(OUT, "demo.pt"))
    print("demo.pt ok", fx["test_py"]["shapes"])


def infer_fixture():
    """BASELINE.json configs[3] at its full shape: 32 content images x 10 random styles, 256x256, gen_state 0
    (test_batch.py:146-164 semantics); reference outputs as 16x average-pooled fingerprints."""
    networks, _ = ref_loader.load()
    cfg = O.config_256_core(gen_state=0, guided=0)
    ga, gb = networks.AdaINGen(3, cfg["gen"]), networks.AdaINGen(3, cfg["gen"])
    ga.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 42, "kaiming"))
    gb.load_state_dict(O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 43, "kaiming"))
    x, _ = seeded_images(1234, 32, 256, 256)
    torch.manual_seed(7)
    styles = torch.randn(10, cfg["gen"]["style_dim"], 1, 1)
    outs = torch.empty(10, 32, 3, 16, 16, dtype=torch.float16)
    with torch.no_grad():
        for b0 in range(0, 32, 8):
            content, _ = ga.encode(x[b0:b0 + 8])
            for j in range(10):
                outs[j, b0:b0 + 8] = _pool(gb.decode(content, styles[j:j + 1].expand(8, -1, -1, -1)), 16)
            print("infer fixture", b0)
    torch.save(dict(seeds=dict(gen_a=42, gen_b=43, img=1234, style=7), styles=styles, outputs=outs, pool=16),
               os.path.join(OUT, "infer_32x10.pt"))


def adam_fixture():
    """torch.optim.Adam (what trainer.py:41-45 instantiates) and the reference ExtraAdam on a toy problem."""
    ref_loader.load()
    ExtraAdam = sys.modules["extraadam"].ExtraAdam
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(257, generator=g)
    grads = [torch.randn(257, generator=g) for _ in range(4)]
    out = {}
    for name, mk in (("adam", lambda p: torch.optim.Adam([p], lr=1e-3, betas=(0.5, 0.999), weight_decay=1e-4)),
                     ("extraadam", lambda p: ExtraAdam([p], lr=1e-3, betas=(0.5, 0.999), weight_decay=1e-4))):
        p = torch.nn.Parameter(p0.clone())
        opt = mk(p)
        hist = []
        for it, gr in enumerate(grads):
            p.grad = gr.clone()
            if name == "extraadam" and it % 2 == 0:
                opt.extrapolation()
            else:
                opt.step()
            hist.append(p.data.clone())
        out[name] = hist
    torch.save(dict(p0=p0, grads=grads, **out), os.path.join(OUT, "adam.pt"))
    print("adam.pt ok")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "adaptation":  # only the fixtures of SURVEY.md 8(f).2
        classifier_fixture()
        step_fixture("g1_adaptation_adam", "adam", 1, 1, 256, 1, 1, adaptation=True)
        sys.exit(0)
    if what == "r2":  # round-2 additions only (benchmark shapes, sampling, demo images)
        sample_fixture()
        demo_fixture()
        autocast_grad_cosines()
        step_fixture("g1_guided_adam_sd8", "adam", 1, 1, 64, 2, 1, style_dim=8)
        step_fixture("g1_guided_adam_b8_256", "adam", 1, 1, 256, 8, 1)
        step_fixture("g1_guided_extraadam_hd512", "extraadam", 1, 1, 512, 1, 2)
        infer_fixture()
        sys.exit(0)
    layers_fixture()
    adam_fixture()
    nets_fixture()
    step_fixture("g1_guided_adam", "adam", 1, 1, 64, 2, 2)
    step_fixture("g0_sampled_extraadam", "extraadam", 0, 0, 64, 1, 2)
    step_fixture("g1_masked_synth_adam", "adam", 1, 1, 64, 2, 1, masked_synth=True)
    classifier_fixture()
    step_fixture("g1_adaptation_adam", "adam", 1, 1, 256, 1, 1, adaptation=True)
    sample_fixture()
    demo_fixture()
    autocast_grad_cosines()
    step_fixture("g1_guided_adam_sd8", "adam", 1, 1, 64, 2, 1, style_dim=8)
    step_fixture("g1_guided_adam_b8_256", "adam", 1, 1, 256, 8, 1)
    step_fixture("g1_guided_extraadam_hd512", "extraadam", 1, 1, 512, 1, 2)
    infer_fixture()
