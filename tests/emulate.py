"""CPU emulation of the tap-GEMM / wgrad descriptor semantics (include/munit_b200.h) -- a CHECKER used
only by tests to validate the host-side planning in munit_b200/geometry.py without a GPU."""
import torch


def _gather_a(flat, rank, dims, strides_b, coords, k0, klen):
    """flat: 1-D float tensor; coords: list of rank-1 LongTensors (dims 1..rank-1) per pixel; returns [P, klen]."""
    P = coords[0].numel()
    valid = torch.ones(P, dtype=torch.bool)
    base = torch.zeros(P, dtype=torch.long)
    for d in range(1, rank):
        cd = coords[d - 1]
        valid &= (cd >= 0) & (cd < dims[d])
        base += cd.clamp(0, dims[d] - 1) * (strides_b[d] // 2)
    ch = k0 + torch.arange(klen)
    chv = (ch >= 0) & (ch < dims[0])
    idx = base[:, None] + ch.clamp(0, dims[0] - 1)[None, :]
    out = flat[idx.clamp(0, flat.numel() - 1)]
    out = out * (valid[:, None] & chv[None, :])
    return out


def tapgemm(plan, a_flat, b_mat, out_flat, bias=None, act="none"):
    p = plan
    n_i, y_i, x_i = torch.meshgrid(torch.arange(p.n_img), torch.arange(p.out_h), torch.arange(p.out_w), indexing="ij")
    n_i, y_i, x_i = n_i.reshape(-1), y_i.reshape(-1), x_i.reshape(-1)
    for ph in range(p.phases):
        acc = torch.zeros(n_i.numel(), p.b_rows, dtype=torch.float32)
        for t in range(p.num_taps):
            off = p.tap_off[t]
            coords = [x_i * p.mx[d] + y_i * p.my[d] + n_i * p.mn[d] + off[d] for d in range(1, p.a_rank)]
            c0 = off[0]  # (+ x0*mx[0].. are zero for all our views)
            A = _gather_a(a_flat, p.a_rank, p.a_dim, p.a_stride, coords, c0, p.chunks * 64)
            k0 = p.b_k0[ph] + t * p.chunks * 64
            acc += A @ b_mat[:, k0:k0 + p.chunks * 64].t()
        if bias is not None:
            acc += bias[None, :]
        if act == "relu":
            acc = acc.relu()
        elif act == "lrelu":
            acc = torch.where(acc > 0, acc, 0.2 * acc)
        elif act == "tanh":
            acc = acc.tanh()
        addr = n_i * p.o_sn + (y_i * p.o_ymul + p.o_yoff[ph]) * p.o_sy + (x_i * p.o_xmul + p.o_xoff[ph]) * p.o_sx
        cols = torch.arange(p.n_store)
        out_flat[addr[:, None] + cols[None, :]] = acc[:, :p.n_store]
    return out_flat


def wgrad(plan, dy_flat, x_flat, dw_flat):
    """(dy, x) as in kernels.wgrad: a swapped plan (tap_on_a) uses x as the A operand and dy as B."""
    p = plan
    a_flat, b_flat = (x_flat, dy_flat) if getattr(p, "tap_on_a", 0) else (dy_flat, x_flat)
    n_i, y_i, x_i = torch.meshgrid(torch.arange(p.n_img), torch.arange(p.out_h), torch.arange(p.out_w), indexing="ij")
    n_i, y_i, x_i = n_i.reshape(-1), y_i.reshape(-1), x_i.reshape(-1)
    m_pad = ((p.m_total + 127) // 128) * 128
    n_pad = ((p.n_total + p.bn - 1) // p.bn) * p.bn
    on_a = getattr(p, "tap_on_a", 0)
    for t in range(p.num_taps):
        off = p.tap_off[t]
        oa = off if on_a else [0] * 5
        ob = [0] * 5 if on_a else off
        ca = [x_i * p.a_mx[d] + y_i * p.a_my[d] + n_i * p.a_mn[d] + oa[d] for d in range(1, p.a_rank)]
        A = _gather_a(a_flat, p.a_rank, p.a_dim, p.a_stride, ca, oa[0], m_pad)  # [P, M]
        cb = [x_i * p.b_mx[d] + y_i * p.b_my[d] + n_i * p.b_mn[d] + ob[d] for d in range(1, p.b_rank)]
        B = _gather_a(b_flat, p.b_rank, p.b_dim, p.b_stride, cb, ob[0], n_pad)  # [P, N]
        G = A.t() @ B  # [M, N]
        m = torch.arange(p.m_total)
        nn_ = torch.arange(p.n_total)
        addr = m[:, None] * p.s_m + t * p.s_t + nn_[None, :] * p.s_n
        dw_flat[addr] += G[:p.m_total, :p.n_total]
    return dw_flat
