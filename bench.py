#!/usr/bin/env python
"""Benchmark of the MUNIT training hot path on B200 (contract: see the task statement / DESIGN.md s6).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # munit_b200 (sm_100a kernels)
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU path (oracle port)

One step = MUNIT_Trainer.dis_update + gen_update (train.py:182-187) on `--batch` image pairs per GPU,
config_256-core (configs/config_256.yaml with semantic/adaptation heads off), 256x256, synthetic
uniform[-1,1] images, seeded kaiming/gaussian weights.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MUNIT train steps/sec (gen+dis) 256^2"
UNIT = "steps/s (1 step = dis_update+gen_update on 8 image pairs)"


def workload_name(batch, hw, hd=False):
    return f"config_{'HD' if hd else '256'}-core train step (dis_update+gen_update), batch {batch}/GPU, {hw}x{hw}"


def load_cfg(hd=False):
    from munit_b200.utils import core_config, get_config

    cfg = core_config(get_config(os.path.join(ROOT, "configs", "config_HD.yaml" if hd else "config_256.yaml")))
    return cfg


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["dram_bytes_read_per_launch"] + d["dram_bytes_write_per_launch"], d["kernel"]
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), tflops_burst=d.get("bf16_tflops"),
                    hbm=d.get("hbm_gbs"), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def synthetic_images(batch, hw, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 3, hw, hw, generator=g) * 2 - 1, torch.rand(batch, 3, hw, hw, generator=g) * 2 - 1)


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(cfg, hw, steps, warmup, batch_equiv, budget_s=150.0):
    """Times dis_update+gen_update of the reference's CPU path at batch 1 (the reference's own batch size,
    config_256.yaml:13).  Uses the unmodified reference when /root/reference is present, else the oracle
    port (oracle/munit_oracle.py, pinned to the reference by tests/golden).  Returns steps/s scaled to
    `batch_equiv` image pairs per step."""
    from oracle import munit_oracle as O
    from oracle import ref_loader

    torch.set_num_threads(os.cpu_count() or 1)
    x_a, x_b = synthetic_images(1, hw, 1234)
    kind = "port"
    if ref_loader.available():
        torch.manual_seed(0)
        tr = ref_loader.make_trainer(cfg)
        kind = "reference"

        def one(it):
            tr.iterations = it
            tr.dis_update(x_a, x_b, cfg)
            tr.gen_update(x_a, x_b, cfg)
    else:
        double = cfg["gen_state"] == 1
        if double:
            gsd = O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 21, cfg["init"])
        else:
            gsd = {k: O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), s, cfg["init"]) for k, s in (("a", 21), ("b", 22))}
        tr = O.OracleTrainer(cfg, gsd, O.init_state_dict(O.dis_spec(cfg["dis"], 3), 23, "gaussian"),
                             O.init_state_dict(O.dis_spec(cfg["dis"], 3), 24, "gaussian"))

        def one(it):
            tr.iterations = it
            tr.dis_update(x_a, x_b)
            tr.gen_update(x_a, x_b)
    t0 = time.time()
    times = []
    for it in range(warmup + steps):
        t1 = time.time()
        one(it)
        if it >= warmup:
            times.append(time.time() - t1)
        if time.time() - t0 > budget_s and times:
            break
    if not times:
        times = [time.time() - t0]
    per_step_b1 = sorted(times)[len(times) // 2]
    value = 1.0 / (per_step_b1 * batch_equiv)
    sample = (f"{len(times)} timed step(s) of dis_update+gen_update at batch 1 (the reference's batch size), {hw}x{hw}, "
              f"fp32, median {per_step_b1:.2f} s/step; steps/s scaled by 1/{batch_equiv} to the batch-{batch_equiv} step")
    return value, kind, sample, per_step_b1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = load_cfg(args.hd)
    hw = cfg["crop_image_height"]
    v, kind, sample, per = cpu_reference_steps(cfg, hw, max(1, min(args.steps, 3)), 1, args.batch)
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1000.0 / v, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=workload_name(args.batch, hw, args.hd), optimizer=cfg["optimizer"]),
                cpu_baseline=dict(value=v, unit=UNIT, cores=os.cpu_count(), kind=kind, sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def profile_tensor_kernels(runner):
    """One extra (untimed) eager step with CUDA events around every tap-GEMM / wgrad launch on the launching
    stream: per-launch device time and direct-form FLOPs of the dominant kernel family."""
    from munit_b200 import kernels as K

    rec = {"tapgemm": [], "wgrad": []}
    orig_t, orig_w = K.tapgemm, K.wgrad

    def wrap(name, fn):
        def inner(plan, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(plan, *a, **k)
            e1.record()
            rec[name].append((e0, e1, getattr(plan, "alg_flops", 0.0), plan))
            return out
        return inner

    K.tapgemm, K.wgrad = wrap("tapgemm", orig_t), wrap("wgrad", orig_w)
    trainer = runner.t
    two = getattr(trainer, "parallel_streams", False)
    try:
        # One stream (no cross-stream overlap inside an event pair) and a GPU that is kept ~0.5 s behind the host,
        # so that every event / kernel is already queued when the device reaches it: the event pairs then bracket
        # device time only, not the Python time spent building the next launch descriptor.
        trainer.parallel_streams = False
        from munit_b200.networks import MsImageDis
        scale_streams, MsImageDis.scale_streams = MsImageDis.scale_streams, False
        runner._prepare_host_state()
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.5 * 1.9e9))
        runner._eager_step()
        runner._advance()
        torch.cuda.synchronize()
    finally:
        trainer.parallel_streams = two
        MsImageDis.scale_streams = scale_streams
        K.tapgemm, K.wgrad = orig_t, orig_w
    out = {}
    detail = []
    for name, lst in rec.items():
        ms = sum(a.elapsed_time(b) for a, b, _, _ in lst)
        fl = sum(f for _, _, f, _ in lst)
        out[name] = dict(launches=len(lst), ms=ms, flops=fl)
        for a, b, f, plan in lst:
            d = dict(kind=name, ms=a.elapsed_time(b), alg_gflop=f / 1e9, n=plan.n_img, oh=plan.out_h, ow=plan.out_w,
                     taps=plan.num_taps, bn=plan.bn)
            if name == "tapgemm":
                d.update(rows=plan.b_rows, chunks=plan.chunks, phases=plan.phases, tile=(plan.tw, plan.th, plan.tn),
                         exec_gflop=plan.flops() / 1e9)
            else:
                d.update(m=plan.m_total, ncols=plan.n_total)
            detail.append(d)
    out["detail"] = detail
    return out


def dominant_launch_time(batch, iters=40):
    """The most frequent launch of the step -- 3x3 256->256 conv forward on [batch, 64, 64] content codes,
    tapgemm_kernel<256,1> -- timed back to back with ONE event pair around `iters` launches on rotating buffers
    (per-launch event pairs add several us to a ~38 us kernel).  Returns (us per launch, TFLOP/s)."""
    from munit_b200 import geometry as G, kernels as K

    n, h, w, c = batch, 64, 64, 256
    hp, wp = h + 2, w + 2
    plan = G.plan_fwd(n, hp, wp, c, 3, 3, 1, 1, c, (h * w * c, w * c, c, 0, 0))
    xs = [torch.randn(n, hp, wp, c, device="cuda").to(torch.bfloat16) for _ in range(4)]
    wt = (torch.randn(c, 9 * c, device="cuda") * 0.02).to(torch.bfloat16)
    ys = [torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda") for _ in range(4)]
    for i in range(4):
        K.tapgemm(plan, xs[i], wt, ys[i])
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.02 * 1.9e9))  # queue everything behind a short spin: no host gaps inside the pair
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        K.tapgemm(plan, xs[i % 4], wt, ys[i % 4])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000.0 / iters
    return us, 2.0 * n * h * w * c * 9 * c / (us * 1e-6) / 1e12


def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from munit_b200 import _lib
    from munit_b200.engine import StepRunner
    from munit_b200.trainer import MUNIT_Trainer

    cfg = load_cfg(args.hd)
    if args.optimizer:
        cfg["optimizer"] = args.optimizer
    hw = args.hw or cfg["crop_image_height"]
    torch.manual_seed(0)  # identical replicas on every rank
    trainer = MUNIT_Trainer(cfg).cuda()
    runner = StepRunner(trainer, cfg, args.batch, hw, use_graph=not args.no_graph, world=world,
                        two_streams=args.two_streams, reuse_forward=bool(args.reuse_forward))
    x_a, x_b = synthetic_images(args.batch, hw, 1234 + rank)
    x_a_h, x_b_h = x_a.pin_memory(), x_b.pin_memory()
    sd = trainer.style_dim
    s_host = [torch.zeros(args.batch, sd, 1, 1).pin_memory() for _ in range(4)]

    def draw_styles():
        # host-generator draws in the reference's order: dis_update (s_a, s_b) then gen_update (s_a, s_b)
        for s in s_host:
            s.copy_(torch.randn(args.batch, sd, 1, 1))

    torch.manual_seed(1000 + rank)
    draw_styles()
    runner.load_inputs(x_a_h, x_b_h, *s_host)
    if args.ncu_step:
        # For `ncu --profile-from-start off ...`: two eager warm-up steps, then exactly one eager step inside the
        # profiler range (about 1800 launches, a couple of minutes under ncu instead of the whole benchmark).
        for _ in range(2):
            runner.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        runner.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_step": "one eager dis_update+gen_update in the profiler range",
                          "launches": runner.launches_per_step}))
        return
    runner.warmup_and_capture(2)
    for _ in range(args.warmup):
        runner.step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- value: inputs resident in HBM
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        runner.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- e2e: host buffers -> H2D every step, loss read back every step
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = None
    for _ in range(args.steps):
        draw_styles()
        runner.load_inputs(x_a_h, x_b_h, *s_host)
        runner.step()
        ls = runner.losses()
        last = (float(ls["loss_dis_total"]), float(ls["loss_gen_total"]))  # D2H + sync
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
    prof = profile_tensor_kernels(runner)
    overlap_note = (sum(len(g.buckets) for g in trainer.grad_sync.values()),
                    sum(g.early_pass for g in trainer.grad_sync.values()), runner.overlap)
    if world > 1:
        runner.release()  # captured graphs reference the communicator: drop them before the process group goes
        dist.barrier()
    # ---- informational: the same step with gen_update picking up dis_update's generator pass (trainer.reuse_forward)
    reuse_line = None
    if world == 1 and not args.reuse_forward and not args.no_graph and cfg["guided"] == 1 and cfg["gen_state"] == 1:
        try:
            runner.release()
            r2 = StepRunner(trainer, cfg, args.batch, hw, use_graph=True, world=1, two_streams=args.two_streams,
                            reuse_forward=True)
            r2.iter = runner.iter
            r2.load_inputs(x_a_h, x_b_h, *s_host)
            r2.warmup_and_capture(1)
            for _ in range(args.warmup):
                r2.step()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(args.steps):
                r2.step()
            g1.record()
            torch.cuda.synchronize()
            ms2 = g0.elapsed_time(g1)
            ls2 = r2.losses()
            reuse_line = dict(value=args.steps / (ms2 / 1000.0), unit=UNIT, ms_per_step=ms2 / args.steps,
                              gpu_launches_per_step=r2.launches_per_step,
                              last_losses=dict(dis=float(ls2["loss_dis_total"]), gen=float(ls2["loss_gen_total"])),
                              note="NOT the headline: same dis_update + gen_update calls with trainer.reuse_forward = True "
                                   "-- gen_update reuses the generator pass (encode x_a, x_b; decode within / across "
                                   "domains) that dis_update ran on the same batch and the same generator weights "
                                   "(guided = 1), 155 of the step's 1395 GMAC per pair are not computed twice; losses, "
                                   "gradients and both optimizer steps as in the headline run "
                                   "(tests/test_trainer_gpu.py::test_forward_reuse_between_updates_is_transparent)")
            r2.release()
            trainer.reuse_forward = False
        except Exception as exc:  # informational line only: never lose the headline over it
            reuse_line = dict(error=repr(exc))
            trainer.reuse_forward = False

    dom_us, dom_tf = dominant_launch_time(args.batch) if rank == 0 else (0.0, 0.0)
    if rank == 0 and args.dump_launches:
        json.dump(prof["detail"], open(args.dump_launches, "w"))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    steps_per_s = args.steps / (ms / 1000.0) * world  # batch-8-equivalent steps, whole job
    e2e_per_s = args.steps / (ms_e2e / 1000.0) * world
    tg = prof["tapgemm"]
    ach = tg["flops"] / (tg["ms"] / 1000.0) / 1e12 if tg["ms"] > 0 else 0.0
    wg = prof["wgrad"]
    ach_w = wg["flops"] / (wg["ms"] / 1000.0) / 1e12 if wg["ms"] > 0 else 0.0
    h2d = 2 * args.batch * 3 * hw * hw * 4 + 4 * args.batch * sd * 4 + 2 * 16
    algo_tflop_step = 2.790 * args.batch * (hw / 256.0) ** 2  # SURVEY.md s8d, per sample pair at 256^2
    cpu_v, cpu_kind, cpu_sample, _ = cpu_reference_steps(cfg, hw, 2, 1, args.batch, budget_s=40.0) \
        if (world == 1 and not args.no_cpu_baseline) else (None, "port", "skipped (N>1 or --no-cpu-baseline)", None)
    line = dict(
        metric=METRIC, value=steps_per_s, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
        data="synthetic",
        config=dict(workload=workload_name(args.batch, hw, args.hd),
                    global_batch=args.batch * world, gen_state=cfg["gen_state"], guided=cfg["guided"],
                    optimizer=cfg["optimizer"], parallelism=f"dp{world}", cuda_graph=not args.no_graph, two_streams=bool(args.two_streams),
                    grad_exchange=("none (1 GPU)" if world == 1 else
                                   ("NCCL all-reduce per ready bucket, overlapped with backward, captured in the step graph "
                                    f"({overlap_note[0]} buckets, {overlap_note[1]} launched before the end "
                                    "of their backward pass)"
                                    if overlap_note[2] else "NCCL all-reduce of the whole arena between three captured segments")),
                    l2="per-step working set (several GB of bf16 activations) >> 126 MB L2; no explicit flush"),
        e2e=dict(value=e2e_per_s, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=8,
                 last_losses=dict(dis=last[0], gen=last[1])),
        gpu_launches=(runner.launches_per_step or 0) * args.steps,
        clocks=clocks,
        roofline=dict(bound="tensor", kernel="tapgemm_kernel<BN> (conv fwd + dgrad, tcgen05)", achieved=ach,
                      peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"] if peaks["tflops"] else None,
                      traffic=profiled_traffic()[0], traffic_kernel=profiled_traffic()[1],
                      peak_source=peaks["src"], launches_per_step=tg["launches"],
                      kernel_ms_per_step=tg["ms"],
                      note="achieved = direct-form FLOPs of all conv fwd/dgrad launches / their device time, one CUDA "
                           "event pair per launch (the pair itself adds several us to each ~40 us launch)",
                      dominant_launch=dict(kernel="tapgemm_kernel<256,1>, 3x3 256->256 forward on [B,64,64]",
                                           us=dom_us, achieved=dom_tf,
                                           frac=dom_tf / peaks["tflops"] if peaks["tflops"] else None,
                                           method="one event pair around 40 back-to-back launches, rotating buffers"),
                      wgrad=dict(achieved=ach_w, launches_per_step=wg["launches"], kernel_ms_per_step=wg["ms"]),
                      step_algorithmic_tflop=algo_tflop_step,
                      step_frac=(algo_tflop_step / (ms / args.steps / 1000.0)) / peaks["tflops"]),
        cpu_baseline=dict(value=cpu_v, unit=UNIT, cores=os.cpu_count(), kind=cpu_kind, sample=cpu_sample),
    )
    if reuse_line is not None:
        line["forward_reuse"] = reuse_line
    if args.reuse_forward:
        line["config"]["reuse_forward"] = True
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_infer(args):
    """test_batch.py semantics (BASELINE.json configs[3]): 32 content images x 10 random style codes at 256^2,
    gen_state 0: one content-encode + 10 (MLP + decoder) passes per step -> output images / s."""
    torch.cuda.set_device(0)
    from munit_b200 import _lib
    from munit_b200.trainer import MUNIT_Trainer

    cfg = load_cfg(args.hd)
    cfg["gen_state"], cfg["guided"] = 0, 0
    hw = args.hw or cfg["crop_image_height"]
    nimg, nstyle = args.batch if args.batch != 8 else 32, 10
    torch.manual_seed(0)
    t = MUNIT_Trainer(cfg).cuda().eval()
    x, _ = synthetic_images(nimg, hw, 1234)
    x_h = x.pin_memory()
    x_d = torch.empty_like(x, device="cuda")
    styles_h = torch.randn(nstyle, t.style_dim, 1, 1).pin_memory()
    styles_d = torch.empty(nstyle, t.style_dim, 1, 1, device="cuda")
    outs = [None] * nstyle

    def step():
        with torch.no_grad():
            c, _ = t.gen_a.encode_act(x_d)
            for j in range(nstyle):
                outs[j] = t.gen_b.decode(c, styles_d[j:j + 1].expand(nimg, -1, -1, -1).contiguous())

    x_d.copy_(x_h); styles_d.copy_(styles_h)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        before = _lib.launches
        step(); step()
        launches = (_lib.launches - before) // 2
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
    for _ in range(max(args.warmup, 3)):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_out = torch.empty(nstyle, nimg, 3, hw, hw).pin_memory()
    f0.record()
    for _ in range(args.steps):
        styles_h.copy_(torch.randn(nstyle, t.style_dim, 1, 1))
        x_d.copy_(x_h, non_blocking=True); styles_d.copy_(styles_h, non_blocking=True)
        g.replay()
        for j in range(nstyle):
            host_out[j].copy_(outs[j], non_blocking=True)
        torch.cuda.synchronize()
    f1.record(); torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1) / args.steps
    peaks = measured_peaks()
    n_out = nimg * nstyle
    tflop = 0.9842 * nimg * (hw / 256.0) ** 2  # 492.1 GMAC per content image with 10 styles
    line = dict(metric="MUNIT style-sampled inference output imgs/sec 256^2", value=n_out / (ms / 1000.0), unit="images/s",
                n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=f"test_batch: {nimg} content images x {nstyle} random styles, {hw}x{hw}, gen_state 0",
                            cuda_graph=True, l2="activations per step >> 126 MB L2; no explicit flush"),
                e2e=dict(value=n_out / (ms_e2e / 1000.0), unit="images/s", h2d_bytes_per_step=int(x.numel() * 4 + styles_h.numel() * 4),
                         d2h_bytes_per_step=int(host_out.numel() * 4)),
                gpu_launches=launches * args.steps,
                roofline=dict(bound="tensor", achieved=tflop / (ms / 1000.0), peak=peaks["tflops"], unit="TFLOP/s",
                              frac=tflop / (ms / 1000.0) / peaks["tflops"], traffic=None, peak_source=peaks["src"],
                              note="whole-step direct-form FLOPs / step time"))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="image pairs per GPU per step")
    ap.add_argument("--hw", type=int, default=0, help="override crop size")
    ap.add_argument("--hd", action="store_true", help="config_HD (512x512, ExtraAdam)")
    ap.add_argument("--optimizer", default="")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--two-streams", type=int, default=1,
                    help="0: one stream; 1: domain-a / domain-b branches on two streams; 2: plus weight gradients on "
                         "companion streams")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reuse-forward", type=int, default=0,
                    help="1: gen_update reuses the generator pass of the preceding dis_update (trainer.reuse_forward); "
                         "the default run reports it next to the headline as `forward_reuse`")
    ap.add_argument("--workload", default="train", choices=["train", "infer"])
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: eager mode, one step inside torch.cuda.profiler.start/stop (use with "
                         "ncu --profile-from-start off); prints no benchmark number")
    ap.add_argument("--dump-launches", default="", help="write per-launch tensor-kernel timings (profile pass) to this json")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.ncu_step:
        args.no_graph, args.two_streams = True, 0
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "infer":
        run_infer(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
