"""Pure-torch emulation of the kernel-side algebra on the PACKED operands (test infrastructure only).

It follows the same steps as the CUDA kernels (hoisted message_layer_1 projections, split message_layer_2
contraction, BatchNorm folded through the sender sum, GEMM + attribute combine for node-level products), so
comparing it with the oracle on CPU validates the host-side packing and the algebraic rewrites without a GPU.
"""
import math

import torch

C_SILU = 1.6791767923989418
C_SIG = 1.8467055342154763
Y1 = 0.4886025119029199


def _unit(v):
    return v / v.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def prep(pos, vel, B, N):
    p = pos.reshape(B, N, 3)
    rel = p[:, None, :, :] - p[:, :, None, :]  # [B, i, j] = pos_j - pos_i
    u = _unit(rel)
    mask = ~torch.eye(N, dtype=torch.bool)
    s = (u * mask[None, :, :, None]).sum(2) / max(N - 1, 1)
    a1 = Y1 * s.reshape(B * N, 3) + Y1 * _unit(vel)
    attr = torch.cat([torch.ones(B * N, 1, dtype=pos.dtype), a1], dim=1)
    x = torch.cat([pos - pos.mean(1, keepdim=True), vel, vel.norm(dim=1, keepdim=True)], dim=1)
    return x, attr


def embed(x, attr, w, bias):
    a0, a1 = attr[:, 0:1], attr[:, 1:4]
    p, v, s = x[:, 0:3], x[:, 3:6], x[:, 6:7]
    pdot, vdot = (p * a1).sum(1, keepdim=True), (v * a1).sum(1, keepdim=True)
    out_s = a0 * w[4] * s + w[2] * pdot + w[3] * vdot + bias
    t = w[5] * s
    out_v = [a1[:, k:k + 1] * t + a0 * (w[0] * p[:, k:k + 1] + w[1] * v[:, k:k + 1]) for k in range(3)]
    return torch.stack([out_s] + out_v, dim=1)  # planar [nodes,4,n]


def node_gemm(x0, x1, w_s, w_v, bias, n_bias):
    x = x0 if x1 is None else torch.cat([x0, x1], dim=2)
    y = torch.stack([x[:, 0] @ w_s] + [x[:, c] @ w_v for c in (1, 2, 3)], dim=1)
    if bias is not None:
        y[:, 0, :n_bias] += bias
    return y


def tp_combine(y, attr, n, gate, bias=None, residual=None, bn_mul=None, bn_add=None):
    n0 = 2 * n if gate else n
    a0, a1 = attr[:, 0:1], attr[:, 1:4]
    z0 = a0 * y[:, 0, :n0] + sum(a1[:, k:k + 1] * y[:, 1 + k, :n0] for k in range(3))
    if bias is not None:
        z0 = z0 + bias
    z1 = [a1[:, k:k + 1] * y[:, 0, n0:] + a0 * y[:, 1 + k, n0:] for k in range(3)]
    if gate:
        s = C_SILU * torch.nn.functional.silu(z0[:, :n])
        g = C_SIG * torch.sigmoid(z0[:, n:])
        out = torch.stack([s] + [g * z for z in z1], dim=1)
    else:
        out = torch.stack([z0] + z1, dim=1)
    if residual is not None:
        out = out + residual
    if bn_mul is not None:
        out = torch.cat([(out[:, 0] * bn_mul[:n] + bn_add).unsqueeze(1), out[:, 1:] * bn_mul[n:]], dim=1)
    return out


def edge_layer(pos, mass, B, N, n, pq, w_edge, w2, bn_mul=None, bn_add=None):
    p = pos.reshape(B, N, 3)
    rel = p[:, None, :, :] - p[:, :, None, :]  # [B,i,j,3] = pos_j - pos_i
    dist = rel.norm(dim=-1, keepdim=True)
    a1 = Y1 * rel / dist.clamp_min(1e-12)
    m = mass.reshape(B, N)
    mm = (m[:, :, None] * m[:, None, :]).unsqueeze(-1)
    pqr = pq.reshape(B, N, 4, 6 * n)
    P, Q = pqr[..., : 3 * n], pqr[..., 3 * n:]
    S = P[:, :, None] + Q[:, None, :]  # [B,i,j,4,3n]
    wd0, wm0, wd1, wm1 = w_edge[:2 * n], w_edge[2 * n:4 * n], w_edge[4 * n:5 * n], w_edge[5 * n:]
    z0 = S[..., 0, :2 * n] + sum(a1[..., k:k + 1] * S[..., 1 + k, :2 * n] for k in range(3)) + dist * wd0 + mm * wm0
    t = S[..., 0, 2 * n:] + dist * wd1 + mm * wm1
    zv = [a1[..., k:k + 1] * t + S[..., 1 + k, 2 * n:] for k in range(3)]
    s1 = C_SILU * torch.nn.functional.silu(z0[..., :n])
    g1 = C_SIG * torch.sigmoid(z0[..., n:])
    v1 = [g1 * z for z in zv]
    dot = sum(a1[..., k:k + 1] * v1[k] for k in range(3))
    y0 = s1 @ w2["ss"] + dot @ w2["vs"] + w2["b"]
    t1 = s1 @ w2["sv"]
    dk = [v1[k] @ w2["vv"] for k in range(3)]
    ms = C_SILU * torch.nn.functional.silu(y0[..., :n])
    gt = C_SIG * torch.sigmoid(y0[..., n:])
    mv = [gt * (a1[..., k:k + 1] * t1 + dk[k]) for k in range(3)]
    mask = (~torch.eye(N, dtype=torch.bool))[None, :, :, None].to(pos.dtype)
    agg = torch.stack([(ms * mask).sum(2)] + [(x * mask).sum(2) for x in mv], dim=2).reshape(B * N, 4, n)
    if bn_mul is not None:
        agg = torch.cat([(agg[:, 0] * bn_mul[:n] + bn_add).unsqueeze(1), agg[:, 1:] * bn_mul[n:]], dim=1)
    return agg


def head(h, attr, w_head):
    a0, a1 = attr[:, 0:1], attr[:, 1:4]
    t = h[:, 0] @ w_head[0]  # [nodes,2]
    d = [h[:, 1 + k] @ w_head[1] for k in range(3)]  # each [nodes,2]
    cols = []
    for o in range(2):
        for k in range(3):
            cols.append(a1[:, k] * t[:, o] + a0[:, 0] * d[k][:, o])
    return torch.stack(cols, dim=1)


def model_forward(packed, n, pos, vel, mass, B, N, return_layers=False):
    """Same kernel sequence as SEGNN.forward_state, on CPU tensors of any float dtype."""
    c = lambda t: None if t is None else t.to(pos.dtype)
    x, attr = prep(pos, vel, B, N)
    h = embed(x, attr, c(packed["embed"]["w"]), c(packed["embed"]["bias"]))
    layers = [h]
    for lw in packed["layers"]:
        m1 = lw["msg1"]
        pq = node_gemm(h, None, c(m1["w_s"]), c(m1["w_v"]), c(m1["bias"]), 2 * n)
        w2 = {k: c(v) for k, v in lw["msg2"].items()}
        agg = edge_layer(pos, mass, B, N, n, pq, c(m1["w_edge"]), w2, c(lw["bn_msg"][0]), c(lw["bn_msg"][1]))
        u1, u2 = lw["upd1"], lw["upd2"]
        g1 = tp_combine(node_gemm(h, agg, c(u1["w_s"]), c(u1["w_v"]), None, 0), attr, n, True, bias=c(u1["bias"]))
        h = tp_combine(node_gemm(g1, None, c(u2["w_s"]), c(u2["w_v"]), None, 0), attr, n, False, bias=c(u2["bias"]),
                       residual=h, bn_mul=c(lw["bn_feat"][0]), bn_add=c(lw["bn_feat"][1]))
        layers.append(h)
    p1 = packed["pool1"]
    hp = tp_combine(node_gemm(h, None, c(p1["w_s"]), c(p1["w_v"]), None, 0), attr, n, True, bias=c(p1["bias"]))
    pred = head(hp, attr, c(packed["head"]))
    return (pred, layers) if return_layers else pred


# ----------------------------------------------------------------------------------------------------------------
# Torch stand-in for the kernel namespace `ops` (test infrastructure): lets tests drive training.forward_train /
# backward_train on CPU in float64. Backward "kernels" are autograd of the forward emulations above, so this checks
# the orchestration and the BatchNorm algebra of training.py, not the CUDA kernels (those are checked on the GPU).
# ----------------------------------------------------------------------------------------------------------------
def edge_messages(pos, mass, B, N, n, p, q, w_edge, w2):
    """Per-edge messages (ms [B,i,j,n], mv 3 x [B,i,j,n]) and the validity mask, as in edge_layer above."""
    pp = pos.reshape(B, N, 3)
    rel = pp[:, None, :, :] - pp[:, :, None, :]
    dist = rel.norm(dim=-1, keepdim=True)
    a1 = Y1 * rel / dist.clamp_min(1e-12)
    m = mass.reshape(B, N)
    mm = (m[:, :, None] * m[:, None, :]).unsqueeze(-1)
    P, Q = p.reshape(B, N, 4, 3 * n), q.reshape(B, N, 4, 3 * n)
    S = P[:, :, None] + Q[:, None, :]
    wd0, wm0, wd1, wm1 = w_edge[:2 * n], w_edge[2 * n:4 * n], w_edge[4 * n:5 * n], w_edge[5 * n:]
    z0 = S[..., 0, :2 * n] + sum(a1[..., k:k + 1] * S[..., 1 + k, :2 * n] for k in range(3)) + dist * wd0 + mm * wm0
    t = S[..., 0, 2 * n:] + dist * wd1 + mm * wm1
    zv = [a1[..., k:k + 1] * t + S[..., 1 + k, 2 * n:] for k in range(3)]
    s1 = C_SILU * torch.nn.functional.silu(z0[..., :n])
    g1 = C_SIG * torch.sigmoid(z0[..., n:])
    v1 = [g1 * z for z in zv]
    dot = sum(a1[..., k:k + 1] * v1[k] for k in range(3))
    y0 = s1 @ w2["ss"] + dot @ w2["vs"] + w2["b"]
    t1 = s1 @ w2["sv"]
    dk = [v1[k] @ w2["vv"] for k in range(3)]
    ms = C_SILU * torch.nn.functional.silu(y0[..., :n])
    gt = C_SIG * torch.sigmoid(y0[..., n:])
    mv = [gt * (a1[..., k:k + 1] * t1 + dk[k]) for k in range(3)]
    mask = (~torch.eye(N, dtype=torch.bool))[None, :, :, None].to(pos.dtype)
    return ms, mv, mask


# e3nn BatchNorm algebra in torch (what segnn_bn_coeffs_fwd / _bwd compute on the device)
def _planar_cols(s, v):
    """per-channel scalar-plane and vector-plane coefficients -> one value per planar column [4n]."""
    return torch.cat([s, v, v, v]).contiguous()


def _bn_forward_coeffs(bn, n, rows, deg, sum_x, sumsq_s, sumsq_v, training):
    """Statistics + folded affine of one e3nn BatchNorm over `rows` rows (rows = E for messages, nodes for
    features). sum_x [n]: sum of the scalar channel over rows; sumsq_s [n]: sum of squares; sumsq_v [n]: sum of
    |v|^2. Returns dict(mean, rs_s, rs_v, mul_s, mul_v, add) where out_s = mul_s * x + add / deg (per row)."""
    w_s, w_v = bn["weight"][:n], bn["weight"][n:]
    if training:
        mean64 = sum_x.double() / rows
        var_s = (sumsq_s.double() / rows - mean64 * mean64).clamp_min(0.0).to(w_s.dtype)
        mean = mean64.to(w_s.dtype)
        var_v = sumsq_v / (3.0 * rows)
    else:
        f = lambda b: b.to(w_s.dtype)
        mean, var_s, var_v = f(bn["running_mean"]), f(bn["running_var"][:n]), f(bn["running_var"][n:])
    rs_s = (var_s + bn["eps"]).rsqrt()
    rs_v = (var_v + bn["eps"]).rsqrt()
    mul_s, mul_v = w_s * rs_s, w_v * rs_v
    add = deg * (bn["bias"] - mean * mul_s)
    return dict(mean=mean, var_s=var_s, var_v=var_v, rs_s=rs_s, rs_v=rs_v, mul_s=mul_s, mul_v=mul_v, add=add)


def _bn_backward_coeffs(bn, st, n, rows, deg, sum_g, sum_gx, training):
    """sum_g [4n] = sum_i G_i, sum_gx [4n] = sum_i G_i * x_i (planar columns; x_i = raw aggregate or pre-norm
    feature). Returns A_s, A_v, B_s, B_v, C_s and the parameter gradients (dweight [2n], dbias [n])."""
    w_s, w_v = bn["weight"][:n], bn["weight"][n:]
    sg_s = deg * sum_g[:n]                                   # sum over rows of dL/dy (scalars)
    sgx_s = sum_gx[:n]                                       # sum over rows of dL/dy * x
    sgx_v = sum_gx[n:2 * n] + sum_gx[2 * n:3 * n] + sum_gx[3 * n:]
    mean, rs_s, rs_v = st["mean"], st["rs_s"], st["rs_v"]
    dgamma_s = rs_s * (sgx_s - mean * sg_s)
    dgamma_v = rs_v * sgx_v
    dbeta = sg_s
    A_s, A_v = w_s * rs_s, w_v * rs_v
    if training:
        c1 = sg_s / rows
        c2 = dgamma_s / rows                                 # mean over rows of dL/dy * x_hat
        B_s = -w_s * rs_s * rs_s * c2
        C_s = -w_s * rs_s * c1 - B_s * mean
        B_v = -w_v * rs_v * rs_v * rs_v * sgx_v / (3.0 * rows)
    else:
        B_s, C_s, B_v = torch.zeros_like(A_s), torch.zeros_like(A_s), torch.zeros_like(A_v)
    return A_s, A_v, B_s, B_v, C_s, torch.cat([dgamma_s, dgamma_v]), dbeta


def _update_running(bn, st):
    mom = bn["momentum"]
    bn["running_mean"].mul_(1 - mom).add_(mom * st["mean"])
    bn["running_var"].mul_(1 - mom).add_(mom * torch.cat([st["var_s"], st["var_v"]]))



class TorchBackend:
    MODE_FP32 = 0

    prep = staticmethod(prep)
    head = staticmethod(lambda h, attr, w_head, n: head(h, attr, w_head))
    embed = staticmethod(lambda x, attr, w, bias, n: embed(x, attr, w, bias))

    @staticmethod
    def node_gemm(x0, x1, w, n_out, bias=None, n_bias=0, split=0, tc=False):
        y = node_gemm(x0, x1, w["w_s"], w["w_v"], bias, n_bias)
        return (y[..., :split].contiguous(), y[..., split:].contiguous()) if split else y

    @staticmethod
    def tp_combine(y, attr, n, gate, bias=None, residual=None, bn_mul=None, bn_add=None):
        return tp_combine(y, attr, n, gate, bias, residual, bn_mul, bn_add)

    @staticmethod
    def edge_layer(mode, pos, mass, B, N, n, p, q, w_edge, w2, bn_mul=None, bn_add=None, want_moments=False):
        ms, mv, mask = edge_messages(pos, mass, B, N, n, p, q, w_edge, w2)
        agg = torch.stack([(ms * mask).sum(2)] + [(x * mask).sum(2) for x in mv], dim=2).reshape(B * N, 4, n)
        mom = torch.cat([(ms * ms * mask).sum(2), (sum(x * x for x in mv) * mask).sum(2)], dim=-1).reshape(B * N, 2 * n)
        assert bn_mul is None
        return (agg, mom) if want_moments else agg

    @staticmethod
    def colsum(x, y=None, mode=0):
        return (x if mode == 0 else (x * x if mode == 1 else x * y)).sum(0)

    @staticmethod
    def lincomb(dy, x, A, B=None, C=None):
        out = A * dy
        if x is not None:
            out = out + B * x
        if C is not None:
            out = out + C
        return out

    @staticmethod
    def add3(a, b, c=None):
        return a + b + (c if c is not None else 0)

    @staticmethod
    def tp_combine_bwd(y, attr, n, gate, bias, dout):
        n0 = 2 * n if gate else n
        with torch.enable_grad():
            yy = y.detach().clone().requires_grad_(True)
            bb = bias.detach().expand(y.shape[0], n0).clone().requires_grad_(True)
            out = tp_combine(yy, attr, n, gate, bias=bb)
            dy, dz0 = torch.autograd.grad(out, [yy, bb], dout)
        return dy, dz0

    @staticmethod
    def node_gemm_wgrad(x0, x1, dy0, dy1, split):
        dy = dy0 if dy1 is None else torch.cat([dy0, dy1], dim=2)
        x = x0 if x1 is None else torch.cat([x0, x1], dim=2)
        dw_s = x[:, 0].t() @ dy[:, 0]
        dw_v = sum(x[:, c].t() @ dy[:, c] for c in (1, 2, 3))
        return dw_s, dw_v

    @staticmethod
    def edge_layer_bwd(pos, mass, B, N, n, p, q, w_edge, w2, bn_a, bn_b, bn_c, dagg):
        with torch.enable_grad():
            leaves = [t.detach().clone().requires_grad_(True) for t in
                      (p, q, w_edge, w2["ss"], w2["vs"], w2["sv"], w2["vv"], w2["b"])]
            pp, qq, we, ss, vs, sv, vv, b = leaves
            ms, mv, mask = edge_messages(pos, mass, B, N, n, pp, qq, we, dict(ss=ss, vs=vs, sv=sv, vv=vv, b=b))
            G = dagg.reshape(B, N, 1, 4, n)
            dms = (bn_a[:n] * G[..., 0, :] + bn_b[:n] * ms + bn_c).detach() * mask
            obj = (dms * ms).sum()
            for k in range(3):
                dmk = (bn_a[n:] * G[..., 1 + k, :] + bn_b[n:] * mv[k]).detach() * mask
                obj = obj + (dmk * mv[k]).sum()
            g = torch.autograd.grad(obj, leaves)
        return g[0], g[1], dict(ss=g[3], vs=g[4], sv=g[5], vv=g[6], b=g[7]), g[2]

    @staticmethod
    def embed_bwd(x_in, attr, dh, n):
        with torch.enable_grad():
            w = torch.zeros(6, n, dtype=dh.dtype, requires_grad=True)
            b = torch.zeros(n, dtype=dh.dtype, requires_grad=True)
            gw, gb = torch.autograd.grad(embed(x_in, attr, w, b), [w, b], dh)
        return gw, gb

    @staticmethod
    def head_bwd(h, attr, w_head, dpred, n):
        with torch.enable_grad():
            hh = h.detach().clone().requires_grad_(True)
            ww = w_head.detach().clone().requires_grad_(True)
            gh, gw = torch.autograd.grad(head(hh, attr, ww), [hh, ww], dpred)
        return gh, gw

    @staticmethod
    def bn_forward_coeffs(bn, n, rows, deg, sums, sq, v_planes, training, update):
        sq_v = sq[n:2 * n] if v_planes == 1 else sq[n:2 * n] + sq[2 * n:3 * n] + sq[3 * n:4 * n]
        st = _bn_forward_coeffs(bn, n, rows, deg, sums[:n], sq[:n], sq_v, training)
        if training and update:
            _update_running(bn, st)
        st["mulcols"] = _planar_cols(st["mul_s"], st["mul_v"])
        st["addcols"] = _planar_cols(st["add"], torch.zeros_like(st["add"]))
        return st

    @staticmethod
    def bn_backward_coeffs(bn, st, n, rows, deg, sum_g, sum_gx, training):
        A_s, A_v, B_s, B_v, C_s, dweight, dbias = _bn_backward_coeffs(bn, st, n, rows, deg, sum_g, sum_gx, training)
        z = torch.zeros_like(C_s)
        return dict(A4=_planar_cols(A_s, A_v), B4=_planar_cols(B_s, B_v), C4=_planar_cols(C_s, z),
                    bn_a=torch.cat([A_s, A_v]), bn_b=torch.cat([B_s, B_v]), bn_c=C_s, dweight=dweight, dbias=dbias)
