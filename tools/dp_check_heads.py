"""Run under torchrun with W ranks: the adaptation head under data parallelism -- synchronised BatchNorm statistics
(all-gather before bn_finalize) and the classifier gradient all-reduce -- against the same updates on one rank with
the whole batch.  256x256 images (the head needs a 64x64 content code), one pair per rank."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from munit_b200 import dp
from munit_b200.trainer import MUNIT_Trainer
from oracle import munit_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = O.config_256_core()
cfg["adaptation"].update(adv_lambda=6, dfeat_lambda=1)
hw, per = 256, 1
gb = per * world
xa, xb = bench.synthetic_images(gb, hw, 77)


def run(world_, rank_, sync):
    torch.manual_seed(0)
    t = MUNIT_Trainer(cfg).cuda()
    for i, c in enumerate((t.domain_classifier_sr_a, t.domain_classifier_sr_b)):
        c.load_state_dict(O.init_classifier_state_dict(40 + i))
    t.cuda()
    x_a, x_b = dp.shard_batch(xa, rank_, world_).cuda(), dp.shard_batch(xb, rank_, world_).cuda()
    t.iterations = 0
    if not sync:  # single process, whole batch: make the collectives no-ops
        from munit_b200 import kernels as K
        K.SYNC_BN = False
    t.domain_classifier_sr_update(x_a, x_b, False, 1.0, 1)
    g = t.classif_opt_sr.g_arena.clone() / (world_ if sync else 1)
    rm = t.domain_classifier_sr_a.BasicBlock1.bn1.running_mean.clone()
    rv = t.domain_classifier_sr_a.BasicBlock2.bn2.running_var.clone()
    loss = t.loss_classifier_sr_update.clone().reshape(1)
    # the fooling loss inside gen_update (classifier frozen, gradient into the encoders)
    torch.manual_seed(5)
    s = [dp.global_style_noise(gb, t.style_dim, rank_, world_).cuda() for _ in range(2)]
    t._gen_backward(x_a, x_b, cfg, None, None, False, s[0], s[1])
    lf = t.loss_classifier_sr.clone().reshape(1)
    torch.cuda.synchronize()
    return g, rm, rv, loss, lf, t.gen_opt.g_arena.clone()


g_dp, rm_dp, rv_dp, l_dp, lf_dp, gg_dp = run(world, rank, True)
dist.all_reduce(l_dp); l_dp /= world
dist.all_reduce(lf_dp); lf_dp /= world
dist.all_reduce(gg_dp); gg_dp /= world
if rank == 0:
    import munit_b200.trainer as T
    _is_init = torch.distributed.is_initialized
    torch.distributed.is_initialized = lambda: False  # the single-rank reference must not all-reduce
    try:
        g_1, rm_1, rv_1, l_1, lf_1, gg_1 = run(1, 0, False)
    finally:
        torch.distributed.is_initialized = _is_init
    cos = lambda a, b: float(torch.dot(a, b) / (a.norm() * b.norm()))
    print("DPHEADS world=%d: classifier loss dp %.6f single %.6f | fooling loss dp %.6f single %.6f" % (
        world, float(l_dp), float(l_1), float(lf_dp), float(lf_1)))
    print("DPHEADS classifier grad cosine %.5f norm ratio %.4f | generator grad cosine %.5f" % (
        cos(g_dp, g_1), float(g_dp.norm() / g_1.norm()), cos(gg_dp, gg_1)))
    print("DPHEADS running_mean max|diff| %.3e (|rm| %.3e)  running_var max|diff| %.3e" % (
        float((rm_dp - rm_1).abs().max()), float(rm_1.abs().max()), float((rv_dp - rv_1).abs().max())))
dist.barrier()
dist.destroy_process_group()
