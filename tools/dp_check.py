"""Run under torchrun with W ranks: one DP step (global batch 2W, 64x64) vs the same step on a single rank with
the whole batch; prints loss and gradient agreement (rank 0)."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from munit_b200 import dp
from munit_b200.engine import StepRunner
from munit_b200.trainer import MUNIT_Trainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = bench.load_cfg()
cfg["guided"] = 0  # exercise the sampled-style path: global draw sliced per rank
hw, per = 64, 2
gb = per * world
xa, xb = bench.synthetic_images(gb, hw, 77)

def run(world_, rank_, batch):
    torch.manual_seed(0)
    t = MUNIT_Trainer(cfg).cuda()
    r = StepRunner(t, cfg, batch, hw, use_graph=False, world=world_)
    torch.manual_seed(123)
    s = [dp.global_style_noise(gb, t.style_dim, rank_, world_) for _ in range(4)]
    r.load_inputs(dp.shard_batch(xa, rank_, world_), dp.shard_batch(xb, rank_, world_), *s)
    r._prepare_host_state(); r._seg_dis()
    if world_ > 1 and not r.overlap: dp.allreduce_arena(t.dis_opt.g_arena)  # (overlap: reduced inside the backward)
    gd = t.dis_opt.g_arena.clone() / world_
    r._seg_mid()
    if world_ > 1 and not r.overlap: dp.allreduce_arena(t.gen_opt.g_arena)
    gg = t.gen_opt.g_arena.clone() / world_
    r._seg_end()
    torch.cuda.synchronize()
    if world_ > 1 and r.overlap and rank_ == 0:
        print("DPCHECK overlap: buckets dis %d gen %d, launched during backward dis %d gen %d" % (
            len(t.grad_sync["dis"].buckets), len(t.grad_sync["gen"].buckets),
            t.grad_sync["dis"].early_pass, t.grad_sync["gen"].early_pass))
    return t, gd, gg

t_dp, gd_dp, gg_dp = run(world, rank, per)
ld = torch.stack([t_dp.loss_dis_total.reshape(()), t_dp.loss_gen_total.reshape(())])
dist.all_reduce(ld); ld /= world
if rank == 0:
    t_1, gd_1, gg_1 = run(1, 0, gb)
    cos = lambda a, b: float(torch.dot(a, b) / (a.norm() * b.norm()))
    print("DPCHECK world=%d: loss_dis dp %.5f single %.5f | loss_gen dp %.5f single %.5f" % (
        world, float(ld[0]), float(t_1.loss_dis_total), float(ld[1]), float(t_1.loss_gen_total)))
    print("DPCHECK grad cosine dis %.5f gen %.5f ; norm ratio dis %.4f gen %.4f" % (
        cos(gd_dp, gd_1), cos(gg_dp, gg_1), float(gd_dp.norm() / gd_1.norm()), float(gg_dp.norm() / gg_1.norm())))
    pc = cos(t_dp.gen_opt.p_arena, t_1.gen_opt.p_arena)
    print("DPCHECK post-step weight max|diff| gen %.3e dis %.3e (lr 1e-4)" % (
        float((t_dp.gen_opt.p_arena - t_1.gen_opt.p_arena).abs().max()), float((t_dp.dis_opt.p_arena - t_1.dis_opt.p_arena).abs().max())))

# captured step: collectives inside the CUDA graph (overlap) vs whole-arena all-reduce between three graphs
def run_graph(overlap, steps=3):
    os.environ["MUNIT_DP_OVERLAP"] = "1" if overlap else "0"
    torch.manual_seed(0)
    t = MUNIT_Trainer(cfg).cuda()
    r = StepRunner(t, cfg, per, hw, use_graph=True, world=world, two_streams=1)
    torch.manual_seed(123)
    s = [dp.global_style_noise(gb, t.style_dim, rank, world) for _ in range(4)]
    r.load_inputs(dp.shard_batch(xa, rank, world), dp.shard_batch(xb, rank, world), *s)
    r.warmup_and_capture(1)
    out = []
    for _ in range(steps):
        r.step()
        out.append((float(t.loss_dis_total), float(t.loss_gen_total)))
    torch.cuda.synchronize()
    return out, t.gen_opt.p_arena.clone(), len(r.graphs[0])

lo, po, ngo = run_graph(True)
ls, ps, ngs = run_graph(False)
# replicas must stay bit-identical: every rank applies the same reduced gradient
chk = po.clone(); dist.broadcast(chk, 0)
flag = torch.tensor([int(torch.equal(chk, po))], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
same = bool(int(flag))
if rank == 0:
    print("DPCHECK graph overlap (%d graph) losses %s" % (ngo, [(round(a, 4), round(b, 4)) for a, b in lo]))
    print("DPCHECK graph split   (%d graphs) losses %s" % (ngs, [(round(a, 4), round(b, 4)) for a, b in ls]))
    print("DPCHECK overlap vs split post-step weights max|diff| %.3e ; replicas identical across ranks: %s" % (
        float((po - ps).abs().max()), same))
dist.barrier()
dist.destroy_process_group()
