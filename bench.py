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
import signal
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------
# Supervisor (one-GPU B200 arm only).  The benchmark proper runs in a child process; if that child makes no progress --
# its own watchdog ends it with exit code 3, or the limit below expires -- it is killed (which tears its CUDA context
# down, hung kernels included) and the headline is measured once more in the conservative configuration (single-CTA
# tap-GEMMs, sequential updates, no extras).  Reason: one multi-GPU run of the full round-2 feature set stopped inside a
# graph replay for a reason that was not isolated (profiles/r2_dp8_hang.md); a single shot of this benchmark must not
# be lost to an event like that.  MUNIT_BENCH_SUPERVISE=0 runs in-process (also automatic under ncu / nsys).
# ------------------------------------------------------------------------------------------------
RETRY_ENV = {"MUNIT_PAIR": "0", "MUNIT_OVERLAP_UPDATES": "0"}


def wants_supervisor(argv, env):
    if env.get("MUNIT_BENCH_CHILD") == "1" or env.get("MUNIT_BENCH_SUPERVISE", "1") == "0":
        return False
    if int(env.get("WORLD_SIZE", "1") or 1) > 1:
        return False  # multi-GPU ranks cannot be restarted one by one: they keep the in-process watchdog only
    if any(k.startswith(("NV_NSIGHT", "NV_COMPUTE_PROFILER", "NSYS_", "CUDA_INJECTION")) for k in env):
        return False  # under a profiler: keep the kernels in the profiled process
    flags = set(argv)
    if flags & {"--ncu-step", "-h", "--help"}:
        return False
    for i, a in enumerate(argv):
        val = a.split("=", 1)[1] if "=" in a else (argv[i + 1] if i + 1 < len(argv) else "")
        if a.split("=", 1)[0] == "--impl" and val == "reference":
            return False
        if a.split("=", 1)[0] == "--workload" and val == "infer":
            return False
        if a.split("=", 1)[0] == "--gpus" and val not in ("", "1"):
            return False
    return True


def supervise(argv, cmd=None, limits=(720.0, 330.0), out=None):
    """Returns the process exit code.  `cmd`: the child command without the benchmark arguments (tests substitute it)."""
    out = out or sys.stdout
    cmd = cmd or [sys.executable, os.path.abspath(__file__)]
    attempts = [({}, [], limits[0]), (RETRY_ENV, ["--no-extras"], limits[1])]
    why = ""
    for i, (env_add, extra_args, limit) in enumerate(attempts):
        env = dict(os.environ, MUNIT_BENCH_CHILD="1")
        env.update(env_add)
        p = subprocess.Popen(cmd + list(argv) + extra_args, env=env, stdout=subprocess.PIPE, text=True)

        def forward(signum, frame, p=p):  # the supervisor is being stopped: take the child along
            p.kill()
            sys.exit(128 + signum)

        old_handlers = {sg: signal.signal(sg, forward) for sg in (signal.SIGTERM, signal.SIGINT)}
        try:
            text, _ = p.communicate(timeout=limit)
            rc = p.returncode
        except subprocess.TimeoutExpired:
            p.kill()
            text, _ = p.communicate()
            rc = -9
        finally:
            for sg, h in old_handlers.items():
                signal.signal(sg, h)
        lines = [ln for ln in (text or "").splitlines() if ln.startswith("{")]
        if rc == 0 and lines:
            line = lines[-1]
            if i > 0:
                try:
                    d = json.loads(line)
                    d["supervisor"] = (f"first attempt ended without a result ({why}); this line is a second run with "
                                       + " ".join(f"{k}={v}" for k, v in RETRY_ENV.items()) + " --no-extras")
                    line = json.dumps(d)
                except ValueError:
                    pass
            out.write(line + "\n")
            out.flush()
            return 0
        why = f"exit code {rc}" if rc != -9 else f"killed after {limit:.0f} s without finishing"
        sys.stderr.write(f"bench.py supervisor: attempt {i + 1} gave no result ({why})\n")
        sys.stderr.flush()
    return 3


if __name__ == "__main__" and wants_supervisor(sys.argv[1:], os.environ):
    sys.exit(supervise(sys.argv[1:]))
if __name__ == "__main__" and os.environ.get("MUNIT_BENCH_CHILD") == "1":
    try:  # die with the supervisor, however it dies (Linux: PR_SET_PDEATHSIG)
        import ctypes
        ctypes.CDLL("libc.so.6", use_errno=True).prctl(1, int(signal.SIGKILL), 0, 0, 0)
    except Exception:
        pass

import torch  # noqa: E402  (after the supervisor: the supervising process never touches CUDA)

METRIC = "MUNIT train steps/sec (gen+dis) 256^2"
UNIT = "steps/s (1 step = dis_update+gen_update on 8 image pairs)"


def workload_name(batch, hw, hd=False):
    return f"config_{'HD' if hd else '256'}-core train step (dis_update+gen_update), batch {batch}/GPU, {hw}x{hw}"


def load_cfg(hd=False):
    """configs/config_256.yaml (or config_HD.yaml) reduced to the benchmark configuration "config_256-core" (SURVEY.md
    D3/D5): semantic / adaptation heads off; config_HD additionally gets `optimizer: extraadam` (BASELINE.json
    configs[4]).  yaml only -- the reference arm must not load the product package (or its shared library)."""
    import yaml

    with open(os.path.join(ROOT, "configs", "config_HD.yaml" if hd else "config_256.yaml")) as f:
        conf = yaml.safe_load(f)
    conf.setdefault("optimizer", "extraadam" if hd else "adam")
    conf["semantic_w"] = 0
    conf["recon_mask"] = 0
    conf["domain_adv_w"] = 0
    conf.setdefault("recon_synth_w", 0)
    ad = dict(conf.get("adaptation") or {})
    for k in ("full_adaptation", "output_classifier_lambda", "output_adv_lambda", "adv_lambda", "dfeat_lambda",
              "sem_seg_lambda"):
        ad[k] = 0
    ad.setdefault("output_classif_freq", 1)
    ad.setdefault("classif_frequency", 15)
    conf["adaptation"] = ad
    return conf


def config_dict(cfg, batch, hw, world, hd=False):
    """The `config` object of the JSON line: the workload only, identical for the B200 arm and the reference arm
    (implementation details of the B200 arm go under `engine`)."""
    return dict(workload=workload_name(batch, hw, hd), global_batch=batch * world, batch_per_gpu=batch, image=f"{hw}x{hw}",
                gen_state=cfg["gen_state"], guided=cfg["guided"], optimizer=cfg["optimizer"], parallelism=f"dp{world}",
                l2="per-step working set (several GB of activations) >> 126 MB L2; no explicit flush")


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
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
def _cpu_trainer(cfg):
    """The reference trainer itself when /root/reference is present (build container), else the oracle port
    (oracle/munit_oracle.py, pinned to the reference by tests/golden).  Returns (one_step(it, x_a, x_b), kind)."""
    from oracle import munit_oracle as O
    from oracle import ref_loader

    if ref_loader.available():
        torch.manual_seed(0)
        tr = ref_loader.make_trainer(cfg)

        def one(it, x_a, x_b):
            tr.iterations = it
            tr.dis_update(x_a, x_b, cfg)
            tr.gen_update(x_a, x_b, cfg)
        return one, "reference"
    double = cfg["gen_state"] == 1
    if double:
        gsd = O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 21, cfg["init"])
    else:
        gsd = {k: O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), s, cfg["init"]) for k, s in (("a", 21), ("b", 22))}
    tr = O.OracleTrainer(cfg, gsd, O.init_state_dict(O.dis_spec(cfg["dis"], 3), 23, "gaussian"),
                         O.init_state_dict(O.dis_spec(cfg["dis"], 3), 24, "gaussian"))

    def one(it, x_a, x_b):
        tr.iterations = it
        tr.dis_update(x_a, x_b)
        tr.gen_update(x_a, x_b)
    return one, "port"


def cpu_reference_steps(cfg, hw, batch, steps, warmup, budget_s):
    """Times REAL dis_update+gen_update steps of the reference's CPU path on `batch` image pairs (the B200 arm's
    step), all host threads, fp32.  Runs `warmup` untimed steps and then up to `steps` timed ones, stopping early
    once `budget_s` is spent (at least one timed step).  Returns (steps/s, kind, sample text, timed, warm)."""
    torch.set_num_threads(os.cpu_count() or 1)
    one, kind = _cpu_trainer(cfg)
    x_a, x_b = synthetic_images(batch, hw, 1234)
    t0 = time.time()
    times = []
    warm = 0
    for it in range(warmup + steps):
        t1 = time.time()
        one(it, x_a, x_b)
        if it >= warmup:
            times.append(time.time() - t1)
        else:
            warm += 1
        if time.time() - t0 > budget_s and times:
            break
    per = sum(times) / len(times)
    sample = (f"{len(times)} timed + {warm} warm-up step(s) of dis_update+gen_update on {batch} image pairs, {hw}x{hw}, "
              f"fp32, {torch.get_num_threads()} threads, mean {per:.2f} s/step (no extrapolation)")
    return 1.0 / per, kind, sample, len(times), warm


def cpu_reference_infer(cfg, hw, nstyle, budget_s):
    """test_batch.py semantics on the host cores: 1 content image x `nstyle` random styles (encode once, decode per
    style; reference scripts/test_batch.py:146-164), output images / s."""
    from oracle import munit_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.init_state_dict(O.gen_spec(cfg["gen"], 3, False), 21, cfg["init"])
    g = O.Gen(sd, cfg["gen"], False)
    x, _ = synthetic_images(1, hw, 1234)
    times = []
    t0 = time.time()
    with torch.no_grad():
        for it in range(4):
            t1 = time.time()
            c, _ = g.encode(x)
            for _ in range(nstyle):
                g.decode(c, torch.randn(1, cfg["gen"]["style_dim"], 1, 1))
            if it:
                times.append(time.time() - t1)
            if time.time() - t0 > budget_s and times:
                break
    per = sum(times) / len(times)
    return nstyle / per, f"{len(times)} timed pass(es) of 1 content image x {nstyle} styles, {hw}x{hw}, fp32, oracle port, {per:.2f} s/pass"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = load_cfg(args.hd)
    hw = args.hw or cfg["crop_image_height"]
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    # a batch-8 step takes ~15 s on 16 cores: bound the run to a few minutes and report what actually ran
    v, kind, sample, timed, warm = cpu_reference_steps(cfg, hw, args.batch, max(1, args.steps), min(args.warmup, 1),
                                                       budget_s=args.ref_budget)
    per = 1000.0 / v          # ms per real step of `batch` pairs
    v = v * args.batch / 8.0  # UNIT counts steps of 8 image pairs
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=world, steps=timed, warmup=warm,
                ms_per_step=per, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=config_dict(cfg, args.batch, hw, world, args.hd),
                cpu_baseline=dict(value=v, unit=UNIT, cores=os.cpu_count(), kind=kind, sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note=f"requested --steps {args.steps} --warmup {args.warmup}; ran {timed} timed / {warm} warm-up "
                     f"steps inside a {args.ref_budget:.0f} s budget (one CPU process; at N > 1 the B200 arm's value is "
                     "the whole job of N GPUs)")
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def _bytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def profile_kernels(runner):
    """One extra (untimed) eager step with CUDA events around every tap-GEMM / wgrad / normalisation launch on the
    launching stream: per-launch device time, direct-form FLOPs of the tensor kernels, bytes of the norm kernels."""
    from munit_b200 import kernels as K

    rec = {"tapgemm": [], "wgrad": []}
    nrec = {k: [] for k in ("norm_stats", "norm_finalize", "norm_apply", "norm_bwd_reduce", "norm_bwd_finalize",
                            "norm_bwd_apply")}
    names_t = ("tapgemm", "wgrad")
    orig = {k: getattr(K, k) for k in names_t + tuple(nrec) + ("norm_finalize_parts", "norm_stats_finalize",
                                                                 "norm_bwd_reduce_finalize")}

    def timed(fn, sink, meta):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            sink.append((e0, e1) + meta(a, k, out))
            return out
        return inner

    K.tapgemm = timed(orig["tapgemm"], rec["tapgemm"], lambda a, k, o: (getattr(a[0], "alg_flops", 0.0), a[0]))
    K.wgrad = timed(orig["wgrad"], rec["wgrad"], lambda a, k, o: (getattr(a[0], "alg_flops", 0.0), a[0]))
    # norm kernels: (elements of the pre-norm tensor, bytes the launch actually touches)
    K.norm_stats = timed(orig["norm_stats"], nrec["norm_stats"], lambda a, k, o: (a[0].numel(), _bytes(a[0])))
    K.norm_finalize = timed(orig["norm_finalize"], nrec["norm_finalize"], lambda a, k, o: (0, _bytes(a[0])))
    K.norm_finalize_parts = timed(orig["norm_finalize_parts"], nrec["norm_finalize"], lambda a, k, o: (0, _bytes(a[0])))
    # fused statistics + finalize / reduce + finalize launches are booked under the pass that reads the big tensors
    K.norm_stats_finalize = timed(orig["norm_stats_finalize"], nrec["norm_stats"], lambda a, k, o: (a[0].numel(), _bytes(a[0])))
    K.norm_bwd_reduce_finalize = timed(orig["norm_bwd_reduce_finalize"], nrec["norm_bwd_reduce"],
                                       lambda a, k, o: (a[3].numel(), _bytes(a[0], a[3])))
    K.norm_apply = timed(orig["norm_apply"], nrec["norm_apply"],
                         lambda a, k, o: (a[0].numel(), _bytes(a[0], o) + (_bytes(a[0]) if a[4] is not None else 0.0)))
    K.norm_bwd_reduce = timed(orig["norm_bwd_reduce"], nrec["norm_bwd_reduce"],
                              lambda a, k, o: (a[3].numel(), _bytes(a[0], a[3])))
    K.norm_bwd_finalize = timed(orig["norm_bwd_finalize"], nrec["norm_bwd_finalize"], lambda a, k, o: (0, _bytes(a[0])))
    K.norm_bwd_apply = timed(orig["norm_bwd_apply"], nrec["norm_bwd_apply"],
                             lambda a, k, o: (a[3].numel(), _bytes(a[0], a[3], o[0], o[1])))
    trainer = runner.t
    two = getattr(trainer, "parallel_streams", False)
    try:
        # One stream (no cross-stream overlap inside an event pair) and a GPU that is kept ~0.5 s behind the host,
        # so that every event / kernel is already queued when the device reaches it: the event pairs then bracket
        # device time only, not the Python time spent building the next launch descriptor.
        trainer.parallel_streams = False
        from munit_b200.networks import MsImageDis
        scale_streams, MsImageDis.scale_streams = MsImageDis.scale_streams, False
        # On the runner's own stream, after one untimed eager step: the caching allocator then owns every block this
        # step needs on this stream (a cudaMalloc in the middle of the step would drain the queue and put host time
        # into the event pairs).
        runner.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(runner.stream):
            runner._prepare_host_state()
            runner._eager_step()
            runner._advance()
            torch.cuda.synchronize()
            for lst in list(rec.values()) + list(nrec.values()):
                lst.clear()
            runner._prepare_host_state()
            torch.cuda._sleep(int(0.5 * 1.9e9))
            runner._eager_step()
            runner._advance()
        torch.cuda.current_stream().wait_stream(runner.stream)
        torch.cuda.synchronize()
    finally:
        trainer.parallel_streams = two
        MsImageDis.scale_streams = scale_streams
        for k, v in orig.items():
            setattr(K, k, v)
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
    norm = {}
    for name, lst in nrec.items():
        ts = [a.elapsed_time(b) for a, b, _, _ in lst]
        # an event pair that straddles a host-side stall (first-touch allocation behind the spin kernel) is not kernel
        # time: pairs more than 20x the median of their family are replaced by the median
        med = sorted(ts)[len(ts) // 2] if ts else 0.0
        ts = [t if t <= 20.0 * med else med for t in ts]
        norm[name] = dict(launches=len(lst), ms=sum(ts), elems=float(sum(e for _, _, e, _ in lst)),
                          bytes=float(sum(by for _, _, _, by in lst)))
    out["norm"] = norm
    return out


def dominant_launch_time(batch, iters=40):
    """The most frequent launch of the step -- 3x3 256->256 conv forward on [batch, 64, 64] content codes,
    tapgemm_pair_kernel<256> (tapgemm_kernel<256> on multi-GPU ranks, engine.StepRunner) -- timed back to back with ONE event pair around `iters` launches on rotating buffers
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


class TrainRun:
    """One timed configuration of the training step on this rank: trainer + StepRunner + pinned host inputs."""

    def __init__(self, cfg, batch, hw, world, rank, args, reuse_forward=False):
        from munit_b200.engine import StepRunner
        from munit_b200.trainer import MUNIT_Trainer

        self.cfg, self.batch, self.hw, self.world, self.rank = cfg, batch, hw, world, rank
        torch.manual_seed(0)  # identical replicas on every rank
        self.trainer = MUNIT_Trainer(cfg).cuda()
        self.runner = StepRunner(self.trainer, cfg, batch, hw, use_graph=not args.no_graph, world=world,
                                 two_streams=args.two_streams, reuse_forward=reuse_forward)
        x_a, x_b = synthetic_images(batch, hw, 1234 + rank)
        self.x_a_h, self.x_b_h = x_a.pin_memory(), x_b.pin_memory()
        self.sd = self.trainer.style_dim
        self.s_host = [torch.zeros(batch, self.sd, 1, 1).pin_memory() for _ in range(4)]
        torch.manual_seed(1000 + rank)
        self.draw_styles()
        self.runner.load_inputs(self.x_a_h, self.x_b_h, *self.s_host)

    def draw_styles(self):
        # host-generator draws in the reference's order: dis_update (s_a, s_b) then gen_update (s_a, s_b)
        for s in self.s_host:
            s.copy_(torch.randn(self.batch, self.sd, 1, 1))

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def prepare(self, warmup):
        self.runner.warmup_and_capture(2)
        for _ in range(warmup):
            self.runner.step()
        torch.cuda.synchronize()

    def time_resident(self, steps):
        """K steps with the inputs resident in HBM; returns ms (max over ranks)."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.runner.step()
        e1.record()
        self.barrier()
        return self._max(e0.elapsed_time(e1))

    def time_e2e(self, steps):
        """K steps through the public entry: per step host randn of the style codes, pinned H2D of images + codes,
        the step, D2H of both total losses."""
        self.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        last = None
        for _ in range(steps):
            self.draw_styles()
            self.runner.load_inputs(self.x_a_h, self.x_b_h, *self.s_host)
            self.runner.step()
            ls = self.runner.losses()
            last = (float(ls["loss_dis_total"]), float(ls["loss_gen_total"]))  # D2H + sync
        f1.record()
        self.barrier()
        return self._max(f0.elapsed_time(f1)), last

    def _max(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt[0])
        return ms

    def h2d_bytes(self):
        return 2 * self.batch * 3 * self.hw * self.hw * 4 + 4 * self.batch * self.sd * 4 + 2 * 16

    def close(self):
        import gc
        self.runner.release()
        self.runner = self.trainer = None
        gc.collect()
        torch.cuda.empty_cache()


def side_train_line(cfg, batch, hw, world, rank, args, steps, warmup, what):
    """Informational extra configuration (never the headline): resident + e2e timing, no profiling pass."""
    run = TrainRun(cfg, batch, hw, world, rank, args)
    try:
        run.prepare(warmup)
        ms = run.time_resident(steps)
        ms_e2e, last = run.time_e2e(steps)
        launches = run.runner.launches_per_step
        peaks = measured_peaks()
        tflop = 2.790 * batch * (hw / 256.0) ** 2  # per GPU per step
        return dict(workload=what, config=config_dict(cfg, batch, hw, world, hw >= 512), steps=steps, warmup=warmup,
                    ms_per_step=ms / steps, steps_per_s=steps / (ms / 1000.0),
                    pairs_per_s=steps * batch * world / (ms / 1000.0),
                    e2e_pairs_per_s=steps * batch * world / (ms_e2e / 1000.0), gpu_launches_per_step=launches,
                    step_frac_of_tensor_peak=(tflop / (ms / steps / 1000.0)) / peaks["tflops"],
                    last_losses=dict(dis=last[0], gen=last[1]))
    finally:
        run.close()


def infer_line(args, steps, with_cpu):
    """test_batch.py semantics (BASELINE.json configs[3]): 32 content images x 10 random style codes at 256^2,
    gen_state 0: one content-encode + 10 (MLP + decoder) passes per step -> output images / s."""
    from munit_b200 import _lib
    from munit_b200.trainer import MUNIT_Trainer

    cfg = load_cfg(False)
    cfg["gen_state"], cfg["guided"] = 0, 0
    hw = cfg["crop_image_height"]
    nimg, nstyle = args.infer_batch, 10
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
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_out = torch.empty(nstyle, nimg, 3, hw, hw).pin_memory()
    f0.record()
    for _ in range(steps):
        styles_h.copy_(torch.randn(nstyle, t.style_dim, 1, 1))
        x_d.copy_(x_h, non_blocking=True); styles_d.copy_(styles_h, non_blocking=True)
        g.replay()
        for j in range(nstyle):
            host_out[j].copy_(outs[j], non_blocking=True)
        torch.cuda.synchronize()
    f1.record(); torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1) / steps
    clocks = sampler.stop()
    peaks = measured_peaks()
    n_out = nimg * nstyle
    tflop = 0.9842 * nimg * (hw / 256.0) ** 2  # 492.1 GMAC per content image with 10 styles
    line = dict(metric="MUNIT style-sampled inference output imgs/sec 256^2", value=n_out / (ms / 1000.0), unit="images/s",
                n_gpus=1, steps=steps, warmup=3, ms_per_step=ms, higher_is_better=True, dtype="bf16", data="synthetic",
                config=dict(workload=f"test_batch: {nimg} content images x {nstyle} random styles, {hw}x{hw}, gen_state 0",
                            cuda_graph=True, l2="activations per step >> 126 MB L2; no explicit flush"),
                e2e=dict(value=n_out / (ms_e2e / 1000.0), unit="images/s", h2d_bytes_per_step=int(x.numel() * 4 + styles_h.numel() * 4),
                         d2h_bytes_per_step=int(host_out.numel() * 4)),
                gpu_launches=launches * steps, clocks=clocks,
                roofline=dict(bound="tensor", achieved=tflop / (ms / 1000.0), peak=peaks["tflops"], unit="TFLOP/s",
                              frac=tflop / (ms / 1000.0) / peaks["tflops"], traffic=None, peak_source=peaks["src"],
                              note="whole-step direct-form FLOPs / step time"))
    del g, t, outs
    import gc
    gc.collect(); torch.cuda.empty_cache()
    if with_cpu:
        v, sample = cpu_reference_infer(cfg, hw, nstyle, budget_s=20.0)
        line["cpu_baseline"] = dict(value=v, unit="images/s", cores=os.cpu_count(), kind="port", sample=sample)
    return line


class Watchdog:
    """Deadline for phases that can stop making progress (a hung collective, a device-side deadlock): when it expires the
    process leaves with os._exit -- after rank 0 has printed the headline line with whatever extras finished (exit 0), or
    with exit code 3 when the headline itself never completed.  Every rank runs its own; no communication needed."""

    def __init__(self):
        import threading
        self.lock = threading.Lock()
        self.deadline, self.what, self.emit = None, "", None
        threading.Thread(target=self._run, daemon=True).start()

    def arm(self, what, seconds, emit=None):
        with self.lock:
            self.what, self.deadline, self.emit = what, time.time() + seconds, emit

    def disarm(self):
        with self.lock:
            self.deadline = None

    def _run(self):
        while True:
            time.sleep(1.0)
            with self.lock:
                d, what, emit = self.deadline, self.what, self.emit
            if d is not None and time.time() > d:
                code = 3
                try:
                    sys.stderr.write(f"bench.py watchdog: '{what}' made no progress before its deadline\n")
                    sys.stderr.flush()
                    if emit is not None:
                        emit(what)
                        code = 0
                finally:
                    os._exit(code)


def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dog = Watchdog()
    dog.arm("headline", 360.0)  # a healthy default run needs ~30 s for the headline (+ ~45 s CPU baseline at N = 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from munit_b200 import _lib  # noqa: F401  (fails loudly when the shared library is missing)

    cfg = load_cfg(args.hd)
    if args.optimizer:
        cfg["optimizer"] = args.optimizer
    hw = args.hw or cfg["crop_image_height"]
    batch = args.batch
    scaling = "weak"
    if args.global_batch:
        assert args.global_batch % world == 0, "--global-batch must be divisible by the number of GPUs"
        batch, scaling = args.global_batch // world, "strong"
    run = TrainRun(cfg, batch, hw, world, rank, args, reuse_forward=bool(args.reuse_forward))
    runner, trainer = run.runner, run.trainer
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
    run.prepare(args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = run.time_resident(args.steps)          # value: inputs resident in HBM
    ms_e2e, last = run.time_e2e(args.steps)     # e2e: host buffers -> H2D every step, losses read back every step
    clocks = sampler.stop() if rank == 0 else None
    prof = profile_kernels(runner)
    overlap_note = (sum(len(g.buckets) for g in trainer.grad_sync.values()),
                    sum(g.early_pass for g in trainer.grad_sync.values()), runner.overlap)
    launches_per_step = runner.launches_per_step or 0
    h2d = run.h2d_bytes()
    run.close()
    del runner, trainer
    if world > 1:
        dist.barrier()
    # ---- the headline line is complete BEFORE any informational extra runs: an extra can fail or even hang (watchdog)
    # without taking the headline down
    dom_us, dom_tf = dominant_launch_time(batch) if rank == 0 else (0.0, 0.0)
    if rank == 0 and args.dump_launches:
        json.dump(prof["detail"], open(args.dump_launches, "w"))
    line = None
    if rank == 0:
        peaks = measured_peaks()
        per8 = batch / 8.0                                            # UNIT counts steps of 8 image pairs
        steps_per_s = args.steps / (ms / 1000.0) * world * per8       # whole job
        e2e_per_s = args.steps / (ms_e2e / 1000.0) * world * per8
        tg = prof["tapgemm"]
        ach = tg["flops"] / (tg["ms"] / 1000.0) / 1e12 if tg["ms"] > 0 else 0.0
        wg = prof["wgrad"]
        ach_w = wg["flops"] / (wg["ms"] / 1000.0) / 1e12 if wg["ms"] > 0 else 0.0
        algo_tflop_step = 2.790 * batch * (hw / 256.0) ** 2  # SURVEY.md s8d, per sample pair at 256^2
        # normalisation family against the HBM roofline.  Algorithmic bytes: forward 4 B per element of
        # the pre-norm tensor (read + write, statistics from the producer), backward 6 B + 4 B for the reduction pass.
        nm = prof["norm"]
        fwd_el = nm["norm_apply"]["elems"]
        bwd_el = nm["norm_bwd_apply"]["elems"]
        norm_ms = sum(v["ms"] for v in nm.values())
        alg_bytes = 4.0 * fwd_el + 10.0 * bwd_el
        moved = sum(v["bytes"] for v in nm.values())
        ach_hbm = alg_bytes / (norm_ms / 1000.0) / 1e9 if norm_ms > 0 else 0.0
        cpu_line = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample="skipped (N>1 or --no-cpu-baseline)")
        if world == 1 and not args.no_cpu_baseline:
            v, kind, sample, _, _ = cpu_reference_steps(cfg, hw, batch, 2, 0, budget_s=25.0)
            cpu_line = dict(value=v * per8, unit=UNIT, cores=os.cpu_count(), kind=kind, sample=sample)
        line = dict(
            metric=METRIC, value=steps_per_s, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms / args.steps, higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="bf16",
            data="synthetic",
            config=config_dict(cfg, batch, hw, world, args.hd),
            engine=dict(cuda_graph=not args.no_graph, two_streams=bool(args.two_streams),
                        grad_exchange=("none (1 GPU)" if world == 1 else
                                       ("NCCL all-reduce per ready bucket, overlapped with backward, captured in the step graph "
                                        f"({overlap_note[0]} buckets, {overlap_note[1]} launched before the end "
                                        "of their backward pass)"
                                        if overlap_note[2] else "NCCL all-reduce of the whole arena between three captured segments"))),
            e2e=dict(value=e2e_per_s, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=8,
                     last_losses=dict(dis=last[0], gen=last[1])),
            gpu_launches=launches_per_step * args.steps,
            clocks=clocks,
            roofline=dict(bound="tensor", kernel="tapgemm_kernel<BN> (conv fwd + dgrad, tcgen05)", achieved=ach,
                          peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"] if peaks["tflops"] else None,
                          traffic=profiled_traffic()[0], traffic_kernel=profiled_traffic()[1],
                          peak_source=peaks["src"], launches_per_step=tg["launches"],
                          kernel_ms_per_step=tg["ms"],
                          note="achieved = direct-form FLOPs of all conv fwd/dgrad launches / their device time inside the "
                               "step, one CUDA event pair per launch (the pair itself adds several us to each ~40 us launch); "
                               "in-step figures are graded against the sustained peak, the isolated launch against burst",
                          dominant_launch=dict(kernel=("tapgemm_pair_kernel<256> (cta_group::2)" if world == 1 else
                                                       "tapgemm_kernel<256>") + ", 3x3 256->256 forward on [B,64,64]",
                                               us=dom_us, achieved=dom_tf, peak=peaks["tflops_burst"],
                                               frac=dom_tf / peaks["tflops_burst"] if peaks["tflops_burst"] else None,
                                               method="one event pair around 40 back-to-back launches, rotating buffers; "
                                                      "denominator = burst bf16 peak (kernel timed alone)"),
                          wgrad=dict(achieved=ach_w, launches_per_step=wg["launches"], kernel_ms_per_step=wg["ms"],
                                     frac=ach_w / peaks["tflops"] if peaks["tflops"] else None),
                          step_algorithmic_tflop=algo_tflop_step,
                          step_frac=(algo_tflop_step / (ms / args.steps / 1000.0)) / peaks["tflops"]),
            roofline_hbm=dict(bound="hbm", kernel="norm_stats / norm_apply / norm_bwd_reduce / norm_bwd_apply (+ finalize)",
                              achieved=ach_hbm, peak=peaks["hbm"], unit="GB/s", frac=ach_hbm / peaks["hbm"] if peaks["hbm"] else None,
                              kernel_ms_per_step=norm_ms, launches_per_step=sum(v["launches"] for v in nm.values()),
                              algorithmic_bytes_per_step=alg_bytes, moved_bytes_per_step=moved,
                              moved_gbs=moved / (norm_ms / 1000.0) / 1e9 if norm_ms > 0 else 0.0,
                              per_kernel={k: dict(ms=v["ms"], launches=v["launches"],
                                                  gbs=(v["bytes"] / (v["ms"] / 1000.0) / 1e9 if v["ms"] > 0 else 0.0))
                                          for k, v in nm.items()},
                              note="achieved = algorithmic bytes (SURVEY.md 8d: 4 B per pre-norm element forward, 10 B "
                                   "backward) / summed device time of the family inside the step; moved = bytes the "
                                   "launches actually touch (statistics pass, 4x up-sampled writes and halos included)"),
            cpu_baseline=cpu_line,
        )
        if args.reuse_forward:
            line["config"]["reuse_forward"] = True
    extras = {}
    side_steps = max(3, min(args.steps, 8 if world == 1 else 4))  # (multi-GPU: short extras, profiles/r2_dp8_hang.md)

    extra_limit = 180.0  # seconds per extra (a healthy one needs 15-60 s)

    def abandon(what):
        if rank == 0:
            extras[what] = dict(error=f"abandoned by the watchdog: no progress within {extra_limit:.0f} s")
            out = dict(line)
            out.update(extras)
            print(json.dumps(out), flush=True)

    extras_deadline = time.time() + 300.0  # all extras together (the supervisor's limit covers headline + this)

    def extra(name, fn):
        left = extras_deadline - time.time()
        if left < 20.0:
            extras[name] = dict(skipped="the time budget of the informational extras was used up")
            return
        dog.arm(name, min(extra_limit, left), abandon)
        try:
            extras[name] = fn()
        except Exception as exc:  # informational lines only: never lose the headline over one of them
            extras[name] = dict(error=repr(exc)[:300])
            import gc
            gc.collect(); torch.cuda.empty_cache()
        finally:
            dog.disarm()

    dog.disarm()

    plain = not (args.hd or args.global_batch or args.no_extras or args.no_graph or args.hw or args.batch != 8)
    # ---- informational: the same step with gen_update picking up dis_update's generator pass (trainer.reuse_forward)
    if plain and world == 1 and not args.reuse_forward and cfg["guided"] == 1 and cfg["gen_state"] == 1:
        def reuse():
            r2 = TrainRun(cfg, batch, hw, world, rank, args, reuse_forward=True)
            try:
                r2.prepare(args.warmup)
                ms2 = r2.time_resident(args.steps)
                ls2 = r2.runner.losses()
                return dict(value=args.steps / (ms2 / 1000.0), unit=UNIT, ms_per_step=ms2 / args.steps,
                            gpu_launches_per_step=r2.runner.launches_per_step,
                            last_losses=dict(dis=float(ls2["loss_dis_total"]), gen=float(ls2["loss_gen_total"])),
                            note="NOT the headline: same dis_update + gen_update calls with trainer.reuse_forward = True "
                                 "-- gen_update reuses the generator pass (encode x_a, x_b; decode within / across "
                                 "domains) that dis_update ran on the same batch and the same generator weights "
                                 "(guided = 1), 155 of the step's 1395 GMAC per pair are not computed twice; losses, "
                                 "gradients and both optimizer steps as in the headline run "
                                 "(tests/test_trainer_gpu.py::test_forward_reuse_between_updates_is_transparent)")
            finally:
                r2.trainer.reuse_forward = False
                r2.close()
        extra("forward_reuse", reuse)
    # ---- informational: BASELINE.json configs[2] (global batch 64 over the N GPUs the driver passed: strong scaling),
    # configs[4] (config_HD, 512^2, ExtraAdam) and configs[3] (32 x 10 style-sampled inference, one GPU)
    if plain and 64 % world == 0:
        extra("global_batch_64", lambda: side_train_line(
            cfg, 64 // world, hw, world, rank, args, side_steps, 3,
            f"config_256-core train step, GLOBAL batch 64 = {64 // world}/GPU x {world} GPU(s) (strong scaling over N)"))
    if plain:
        cfg_hd = load_cfg(True)
        extra("hd", lambda: side_train_line(cfg_hd, args.hd_batch, cfg_hd["crop_image_height"], world, rank, args,
                                            side_steps + side_steps % 2, 4,
                                            f"config_HD train step (ExtraAdam, extrapolation/step pairs), "
                                            f"batch {args.hd_batch}/GPU, 512x512"))
    if plain and world == 1:
        extra("infer", lambda: infer_line(args, side_steps, not args.no_cpu_baseline))
    if rank == 0:
        line.update(extras)
        print(json.dumps(line), flush=True)
    dog.arm("teardown", 60.0, lambda what: None)  # (the line is out: leave with exit code 0 whatever happens now)
    if world > 1:
        dist.destroy_process_group()
    dog.disarm()


def run_infer(args):
    torch.cuda.set_device(0)
    line = infer_line(args, args.steps, not args.no_cpu_baseline)
    line.update(scaling="weak", vs_baseline=None)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="image pairs per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: total image pairs per step, split over the GPUs (BASELINE.json configs[2]: 64)")
    ap.add_argument("--hw", type=int, default=0, help="override crop size")
    ap.add_argument("--hd", action="store_true", help="config_HD (512x512, ExtraAdam)")
    ap.add_argument("--hd-batch", type=int, default=8, help="image pairs per GPU of the informational config_HD line")
    ap.add_argument("--infer-batch", type=int, default=32, help="content images of the inference workload")
    ap.add_argument("--optimizer", default="")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--two-streams", type=int, default=1,
                    help="0: one stream; 1: domain-a / domain-b branches on two streams; 2: plus weight gradients on "
                         "companion streams")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only (skip forward_reuse / global_batch_64 / hd / infer)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds of CPU time the reference arm may spend")
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
