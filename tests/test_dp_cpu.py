"""Data-parallel host logic on the CPU with world_size 2 (gloo): bucketed all-reduce of a flat gradient arena,
batch sharding, bit-exact global style codes, and the claim the DP design rests on -- per-rank gradients of
the MUNIT losses on batch shards, summed and scaled by 1/W, equal the gradient of the global batch (checked
with the CPU oracle, which is the reference's arithmetic)."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, initfile, result_q):
    sys.path.insert(0, ROOT)
    from munit_b200 import dp
    from oracle import munit_oracle as O

    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        # 1. bucketed all-reduce == plain sum
        g = torch.Generator().manual_seed(rank)
        arena = torch.randn(100_003, generator=g)
        ref = arena.clone()
        dist.all_reduce(ref)
        dp.allreduce_arena(arena, bucket_bytes=64 * 1024)
        assert torch.equal(arena, ref)
        # 2. bit-exact style codes: global draw sliced per rank
        torch.manual_seed(5)
        mine = dp.global_style_noise(4, 16, rank, world)
        torch.manual_seed(5)
        full = torch.randn(4, 16, 1, 1)
        assert torch.equal(mine, full[rank * 2:(rank + 1) * 2])
        # 3. DP gradient == global-batch gradient (discriminator loss on a tiny net keeps this fast)
        cfg = O.config_256_core()
        dsd = O.init_state_dict(O.dis_spec(cfg["dis"], 3), 3, "gaussian")
        params = {k: v.clone().requires_grad_(True) for k, v in dsd.items()}
        gi = torch.Generator().manual_seed(11)
        fake, real = torch.rand(2, 3, 64, 64, generator=gi) * 2 - 1, torch.rand(2, 3, 64, 64, generator=gi) * 2 - 1
        loss_full = O.calc_dis_loss(params, cfg["dis"], fake, real)
        g_full = torch.autograd.grad(loss_full, list(params.values()))
        loss_r = O.calc_dis_loss(params, cfg["dis"], dp.shard_batch(fake, rank, world), dp.shard_batch(real, rank, world))
        g_r = torch.autograd.grad(loss_r, list(params.values()))
        flat = torch.cat([t.reshape(-1) for t in g_r])
        dp.allreduce_arena(flat, bucket_bytes=1 << 20)
        flat /= world
        ref_flat = torch.cat([t.reshape(-1) for t in g_full])
        err = float((flat - ref_flat).norm() / ref_flat.norm())
        assert err < 1e-5, err
        # 4. readiness-driven bucketed all-reduce (dp.GradSync): buckets are whole parameters cut from the tail, a
        #    bucket is reduced as soon as every use of every parameter in it is done, always in the same order
        sizes = [64, 640, 128, 1280, 64, 320, 192]
        offs, o = [], 0
        for n in sizes:
            offs.append((o, n))
            o += n
        g = torch.Generator().manual_seed(100 + rank)
        arena = torch.randn(o, generator=g)
        ref = arena.clone()
        dist.all_reduce(ref)
        gs = dp.GradSync(arena, offs, bucket_bytes=4 * 700)
        assert gs.buckets[0][1] == o and gs.buckets[-1][0] == 0
        assert all(a[0] == b[1] for a, b in zip(gs.buckets, gs.buckets[1:]))
        starts = {s for s, _ in offs}
        assert all(s in starts for s, _ in gs.buckets)  # parameters are never split
        views = [arena[s:s + n] for s, n in offs]
        gs.begin()
        for v in views[:-1]:  # the last parameter is not used in this pass: its gradient (zero) is already final
            gs.note_use(v)
        gs.note_use(views[1])  # one parameter used twice (a weight shared by two calls)
        order = [5, 4, 3, 1, 2, 0, 1] if rank == 0 else [4, 5, 1, 3, 2, 1, 0]  # ranks may finish in different orders
        seen_early = []
        for i in order:
            gs.note_done(views[i])
            seen_early.append(gs.launched_early)
        # buckets: [3,4,5,6] [1,2] [0].  Rank 0 completes bucket [0] before bucket [1,2]: it is held back so that
        # every rank issues the collectives in the same order
        assert seen_early == ([0, 0, 1, 1, 1, 1, 3] if rank == 0 else [0, 0, 0, 1, 1, 2, 3]), seen_early
        gs.finish()
        assert torch.equal(arena, ref)
        # same again with every parameter used: buckets go out during the "backward"
        arena.copy_(torch.randn(o, generator=g))
        ref = arena.clone()
        dist.all_reduce(ref)
        gs.begin()
        for v in views:
            gs.note_use(v)
        for i in reversed(range(len(views))):
            gs.note_done(views[i])
        assert gs.launched_early == 3 + len(gs.buckets)  # (the counter is cumulative)
        gs.finish()
        assert torch.equal(arena, ref)
        foreign = torch.zeros(8)
        gs.begin(); gs.note_use(foreign); gs.note_done(foreign); gs.finish()  # buffers of other arenas are ignored
        result_q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        result_q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_dp_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        initfile = os.path.join(d, "init")
        procs = [ctx.Process(target=_worker, args=(r, world, initfile, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join(60)
    assert all(r[1] == "ok" for r in res), res


def test_bucket_slices_cover_arena():
    from munit_b200 import dp

    for n in (1, 7, 1024, 100_003):
        sl = dp.bucket_slices(n, 4096)
        assert sl[0][1] == n and sl[-1][0] == 0
        assert all(a[0] == b[1] for a, b in zip(sl, sl[1:]))
        assert sum(e - s for s, e in sl) == n
