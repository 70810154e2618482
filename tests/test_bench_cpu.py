"""bench.py host logic that needs no GPU: the watchdog that keeps a hung informational extra (or a hung headline) from
turning into a silent 10-minute NCCL timeout."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code):
    return subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, timeout=300, text=True)


def test_watchdog_emits_the_headline_and_exits_zero():
    r = _run("import bench, time, json\n"
             "d = bench.Watchdog()\n"
             "d.arm('hd', 1.0, lambda w: print(json.dumps({'value': 1.5, w: {'error': 'abandoned'}}), flush=True))\n"
             "time.sleep(60)\n")
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.5 and "error" in line["hd"]
    assert "watchdog" in r.stderr


def test_watchdog_without_a_headline_fails_loudly_and_disarm_works():
    r = _run("import bench, time\nd = bench.Watchdog()\nd.arm('headline', 1.0)\ntime.sleep(60)\n")
    assert r.returncode == 3 and r.stdout.strip() == ""
    r = _run("import bench, time\nd = bench.Watchdog()\nd.arm('x', 1.0)\nd.disarm()\ntime.sleep(3)\nprint('alive')\n")
    assert r.returncode == 0 and r.stdout.strip() == "alive"


def test_run_b200_assembles_and_prints_one_line_with_everything_mocked(monkeypatch, capsys):
    """Control flow of the B200 arm without a GPU: trainer / runner / kernels replaced by stand-ins; the JSON line must
    carry the contract keys, the extras, and be printed exactly once -- also when an extra raises."""
    import argparse
    import types

    import torch

    import bench

    class FakeRun:
        def __init__(self, cfg, batch, hw, world, rank, args, reuse_forward=False):
            gs = types.SimpleNamespace(buckets=[1, 2], early_pass=2)
            self.trainer = types.SimpleNamespace(grad_sync={"gen": gs}, reuse_forward=False)
            self.runner = types.SimpleNamespace(launches_per_step=1400, overlap=False,
                                                losses=lambda: dict(loss_dis_total=1.0, loss_gen_total=2.0))
        def prepare(self, warmup): pass
        def time_resident(self, steps): return 35.0 * steps
        def time_e2e(self, steps): return 36.0 * steps, (9.0, 40.0)
        def h2d_bytes(self): return 123
        def close(self): pass

    fam = dict(launches=1, ms=1.0, flops=1e12)
    nrm = {k: dict(launches=1, ms=1.0, elems=1e6, bytes=4e6) for k in
           ("norm_stats", "norm_finalize", "norm_apply", "norm_bwd_reduce", "norm_bwd_finalize", "norm_bwd_apply")}
    monkeypatch.setattr(bench, "TrainRun", FakeRun)
    monkeypatch.setattr(bench, "profile_kernels", lambda r: dict(tapgemm=fam, wgrad=fam, norm=nrm, detail=[]))
    monkeypatch.setattr(bench, "dominant_launch_time", lambda b: (36.0, 1050.0))
    monkeypatch.setattr(bench, "cpu_reference_steps", lambda *a, **k: (0.05, "port", "mock", 2, 0))
    monkeypatch.setattr(bench, "infer_line", lambda *a, **k: dict(value=8000.0))
    calls = []

    def side(cfg, batch, hw, world, rank, args, steps, warmup, what):
        calls.append(what)
        if "config_HD" in what:
            raise RuntimeError("boom")
        return dict(ms_per_step=100.0)

    monkeypatch.setattr(bench, "side_train_line", side)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)

    class FakeSampler:
        def __init__(self, i): pass
        def start(self): pass
        def stop(self): return dict(sm_mhz=1900, sm_max_mhz=1965, reasons=[])

    monkeypatch.setattr(bench, "ClockSampler", FakeSampler)
    import munit_b200
    monkeypatch.setitem(sys.modules, "munit_b200._lib", types.SimpleNamespace())
    monkeypatch.setattr(munit_b200, "_lib", types.SimpleNamespace(), raising=False)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(gpus=1, steps=20, warmup=3, impl="b200", batch=8, global_batch=0, hw=0, hd=False, hd_batch=8,
                              infer_batch=32, optimizer="", no_graph=False, two_streams=1, no_cpu_baseline=False,
                              no_extras=False, ref_budget=150.0, reuse_forward=0, workload="train", ncu_step=False,
                              dump_launches="")
    bench.run_b200(args)
    out = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(out) == 1
    line = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert abs(line["ms_per_step"] - 35.0) < 1e-9 and abs(line["value"] - 1000.0 / 35.0) < 1e-6
    assert line["roofline"]["frac"] and line["roofline_hbm"]["frac"] and line["cpu_baseline"]["kind"] == "port"
    assert line["forward_reuse"]["value"] and line["global_batch_64"]["ms_per_step"] == 100.0
    assert "boom" in line["hd"]["error"] and line["infer"]["value"] == 8000.0
    assert len(calls) == 2


def test_gather_segment_table_layout():
    """kernels.gather_seg_table (host side of munit_gather_cast_multi): five 64-bit words per segment -- source, index
    map (0 = identity), destination, element count, first block -- with block ranges that tile the launch grid."""
    import torch

    from munit_b200 import kernels as K

    src = torch.zeros(10000)
    segs = [(src[:100], torch.zeros(5000, dtype=torch.int32), torch.zeros(5000, dtype=torch.bfloat16)),
            (src[100:2148], None, torch.zeros(2048, dtype=torch.bfloat16)),
            (src[3000:], torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.bfloat16))]
    table, nseg, nblocks = K.gather_seg_table(segs, "cpu")
    assert nseg == 3 and table.dtype == torch.int64 and tuple(table.shape) == (3, 5)
    per = K.GATHER_BLOCK
    want_blocks = [(5000 + per - 1) // per, 1, 1]
    assert nblocks == sum(want_blocks)
    b0 = 0
    for row, (s, i, d), nb in zip(table.tolist(), segs, want_blocks):
        assert row == [s.data_ptr(), 0 if i is None else i.data_ptr(), d.data_ptr(), d.numel(), b0]
        b0 += nb


def test_supervisor_passes_the_line_through_and_retries_a_hung_child(tmp_path):
    """bench.supervise: a healthy child's JSON line goes out unchanged; a child that hangs (or dies) is killed and the
    benchmark runs once more with the conservative environment, the line then says so; two failures -> exit code 3."""
    import io

    import bench

    fake = tmp_path / "fake_bench.py"
    fake.write_text(
        "import json, os, sys, time\n"
        "mode = os.environ.get('FAKE_MODE', 'ok')\n"
        "assert os.environ.get('MUNIT_BENCH_CHILD') == '1'\n"
        "conservative = os.environ.get('MUNIT_PAIR') == '0' and '--no-extras' in sys.argv\n"
        "if mode == 'hang_first' and not conservative:\n"
        "    time.sleep(600)\n"
        "if mode == 'die_first' and not conservative:\n"
        "    sys.exit(3)\n"
        "if mode == 'always_die':\n"
        "    sys.exit(1)\n"
        "print('some log line')\n"
        "print(json.dumps({'metric': 'm', 'value': 28.7 if not conservative else 27.0, 'args': sys.argv[1:]}))\n")
    cmd = [sys.executable, str(fake)]

    def run(mode, limits=(5.0, 30.0)):
        os.environ["FAKE_MODE"] = mode
        try:
            buf = io.StringIO()
            rc = bench.supervise(["--steps", "20"], cmd=cmd, limits=limits, out=buf)
        finally:
            del os.environ["FAKE_MODE"]
        return rc, buf.getvalue()

    rc, text = run("ok")
    assert rc == 0 and text.count("\n") == 1
    d = json.loads(text)
    assert d["value"] == 28.7 and d["args"] == ["--steps", "20"] and "supervisor" not in d
    for mode in ("hang_first", "die_first"):
        rc, text = run(mode)
        d = json.loads(text)
        assert rc == 0 and d["value"] == 27.0 and d["args"] == ["--steps", "20", "--no-extras"]
        assert "MUNIT_PAIR=0" in d["supervisor"] and "first attempt" in d["supervisor"]
    rc, text = run("always_die")
    assert rc == 3 and text == ""


def test_supervisor_is_only_for_the_one_gpu_b200_arm():
    import bench

    w = bench.wants_supervisor
    assert w([], {}) and w(["--gpus", "1", "--steps", "20", "--warmup", "3"], {}) and w(["--gpus=1"], {})
    assert not w(["--gpus", "8"], {}) and not w([], {"WORLD_SIZE": "2"}) and not w(["--impl", "reference"], {})
    assert not w(["--impl=reference"], {}) and not w(["--workload", "infer"], {}) and not w(["--ncu-step"], {})
    assert not w([], {"MUNIT_BENCH_CHILD": "1"}) and not w([], {"MUNIT_BENCH_SUPERVISE": "0"})
    assert not w([], {"NV_COMPUTE_PROFILER_PERFWORKS_DIR": "/x"}) and w([], {"WORLD_SIZE": "1", "RANK": "0"})
