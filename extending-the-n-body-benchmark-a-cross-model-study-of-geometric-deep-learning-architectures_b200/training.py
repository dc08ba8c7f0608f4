"""Forward with train-mode BatchNorm and the hand-written backward of the SEGNN kernel sequence
(trainer.py:233-358 `train_one_step`: forward -> loss -> backward; models/segnn/segnn.py:150-304).

Nothing per-edge is saved by the forward: the edge-layer backward (K3^T) recomputes the messages from the saved
node-level projections. Per layer the saved tensors are all [nodes, .]: the layer input, the hoisted projections
P/Q, the raw aggregate (sum over senders of the un-normalised messages), the two node GEMM outputs and the
pre-BatchNorm features.

e3nn BatchNorm (segnn.py:233-235) in train mode is an affine map given its batch statistics; for the message
BatchNorm those statistics are sums over all E edges, which the edge kernel delivers as per-receiver partial sums
(sum_j m, sum_j m^2) reduced here by a deterministic column sum. With G_i = dL/d agg_i, the gradient that reaches
every message of receiver i is dm_ij = A * G_i + B * m_ij + C with per-channel A, B, C computed from node-level
reductions (sum_i G_i, sum_i G_i * agg_raw_i): `segnn_bn_coeffs_bwd` (the algebra is restated in torch in
tests/emulate.py and checked against autograd through the oracle).

`backend` is the kernel namespace (``ops``: the C ABI). Tests substitute a torch emulation to check this
orchestration and the BatchNorm algebra on CPU; the product path always uses ``ops``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops as _ops

__all__ = ["forward_train", "backward_train", "SegnnTrainFunction", "SegnnTrainFunctionFlat", "build_pack_map",
           "flatten_packed", "unflatten_packed", "attach_bn_buffers"]


# ---------------------------------------------------------------------------------------------------------------
# packed-weight tree <-> flat tuple (autograd.Function inputs must be a flat list of tensors)
# ---------------------------------------------------------------------------------------------------------------
def flatten_packed(tree):
    """Deterministic depth-first flattening (dict keys sorted). Returns (leaves, spec)."""
    leaves: List[torch.Tensor] = []

    def rec(node):
        if isinstance(node, dict):
            return {"d": [(k, rec(node[k])) for k in sorted(node.keys())]}
        if isinstance(node, (list, tuple)):
            return {"l": [rec(v) for v in node]}
        if node is None:
            return {"n": None}
        leaves.append(node)
        return {"t": len(leaves) - 1}

    spec = rec(tree)
    return leaves, spec


def unflatten_packed(leaves, spec):
    if "d" in spec:
        return {k: unflatten_packed(leaves, s) for k, s in spec["d"]}
    if "l" in spec:
        return [unflatten_packed(leaves, s) for s in spec["l"]]
    if "n" in spec:
        return None
    return leaves[spec["t"]]


def _colsum2_of(be):
    """The paired column sum of a kernel namespace; test back ends that only emulate `colsum` get two calls."""
    fn = getattr(be, "colsum2", None)
    if fn is not None:
        return fn
    return lambda xa, ya, ma, xb, yb, mb: (be.colsum(xa, ya, ma), be.colsum(xb, yb, mb))


# ---------------------------------------------------------------------------------------------------------------
# forward (fp32 kernels), saving node-level tensors only
# ---------------------------------------------------------------------------------------------------------------
def forward_train(W: Dict, n: int, pos, vel, mass, B: int, N: int, bn_training: bool, backend=None,
                  update_running_stats: bool = True, keep_rows: bool = False):
    """W: packed weights (see SEGNN.packed_train). Returns (pred [nodes,6], saved).  keep_rows (a backward pass will
    follow): large graphs, whose edge layers run in GEMM form, leave their edge rows in HBM for the backward call when
    they fit ops.GEMM_FORM_KEEP_BYTES_PER_LAYER; everything else stores node-level tensors only."""
    be = backend or _ops
    colsum2 = _colsum2_of(be)
    nodes, deg = B * N, N - 1
    E = nodes * deg
    x_in, attr = be.prep(pos, vel, B, N)
    h = be.embed(x_in, attr, W["embed"]["w"], W["embed"]["bias"], n)
    saved = dict(x_in=x_in, attr=attr, layers=[], n=n, B=B, N=N, pos=pos, mass=mass, bn_training=bn_training)
    for lw in W["layers"]:
        m1, m2, u1, u2 = lw["msg1"], lw["msg2"], lw["upd1"], lw["upd2"]
        p, q = be.node_gemm(h, None, m1, 6 * n, bias=m1["bias"], n_bias=2 * n, split=3 * n)
        rows = None
        if keep_rows and backend is None and _ops.gemm_form_keeps_rows(B, N, n):
            agg_raw, mom, rows = _ops.edge_layer_gemm_fwd(pos, mass, B, N, n, p, q, m1["w_edge"], m2, None, None,
                                                          want_moments=True, keep_rows=True)
        else:
            agg_raw, mom = be.edge_layer(be.MODE_FP32, pos, mass, B, N, n, p, q, m1["w_edge"], m2, None, None,
                                         want_moments=True)
        rec = dict(h=h, p=p, q=q, agg_raw=agg_raw, rows=rows)
        if lw["bn_msg"] is not None:
            sums, sq = colsum2(agg_raw.view(nodes, 4 * n), None, 0, mom, None, 0)
            st = be.bn_forward_coeffs(lw["bn_msg"], n, float(E), float(deg), sums, sq, 1, bn_training,
                                      update_running_stats)
            agg = be.lincomb(agg_raw.view(nodes, 4 * n), None, st["mulcols"], None, st["addcols"]).view(nodes, 4, n)
            rec["bn_msg"] = st
        else:
            agg = agg_raw
        y1 = be.node_gemm(h, agg, u1, 3 * n)
        g1 = be.tp_combine(y1, attr, n, True, bias=u1["bias"])
        y2 = be.node_gemm(g1, None, u2, 2 * n)
        pre = be.tp_combine(y2, attr, n, False, bias=u2["bias"], residual=h)
        rec.update(agg=agg, y1=y1, g1=g1, y2=y2, pre=pre)
        if lw["bn_feat"] is not None:
            flat = pre.view(nodes, 4 * n)
            sums, sq = colsum2(flat, None, 0, flat, None, 1)
            st = be.bn_forward_coeffs(lw["bn_feat"], n, float(nodes), 1.0, sums, sq, 3, bn_training,
                                      update_running_stats)
            h = be.lincomb(flat, None, st["mulcols"], None, st["addcols"]).view(nodes, 4, n)
            rec["bn_feat"] = st
        else:
            h = pre
        saved["layers"].append(rec)
    p1 = W["pool1"]
    yp = be.node_gemm(h, None, p1, 3 * n)
    hp = be.tp_combine(yp, attr, n, True, bias=p1["bias"])
    pred = be.head(hp, attr, W["head"], n)
    saved.update(h_last=h, yp=yp, hp=hp)
    return pred, saved


# ---------------------------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------------------------
def _transposed(w):
    if "w_s_t" in w:  # gathered together with the other operand blocks (build_pack_map)
        return dict(w_s=w["w_s_t"], w_v=w["w_v_t"])
    return dict(w_s=w["w_s"].t().contiguous(), w_v=w["w_v"].t().contiguous())


class _SideWork:
    """Weight-gradient work (wgrad GEMMs, bias column sums) is off the critical dgrad chain: on CUDA it goes to a side
    stream (ops.side_stream) after a fork on the current stream and is joined once at the end of the backward pass."""

    def __init__(self, ref):
        self.stream = _ops.side_stream(ref.device, 1) if ref.is_cuda else None
        self.used = False
        self.keep = []  # closures (and through them the tensors the side kernels read) stay alive until join(): the
        #                 caching allocator must not hand their memory to a later main-stream allocation meanwhile

    def run(self, fn):
        if self.stream is None:
            return fn()
        self.stream.wait_stream(torch.cuda.current_stream(self.stream.device))
        self.used = True
        self.keep.append(fn)
        with torch.cuda.stream(self.stream):
            return fn()

    def join(self):
        if self.stream is not None and self.used:
            torch.cuda.current_stream(self.stream.device).wait_stream(self.stream)
        self.keep.clear()


def _node_tp_backward(be, w, x0, x1, y, attr, n, gate, dout, split_out: Optional[int] = None, side=None):
    """Backward of node_gemm(x0|x1; w) -> tp_combine(gate, bias). Returns (dx (tensor or pair), grads dict)."""
    nodes = y.shape[0]
    n0 = 2 * n if gate else n
    dy, dz0 = be.tp_combine_bwd(y, attr, n, gate, w["bias"], dout)

    def weight_side():
        dbias = be.colsum(dz0.view(nodes, n0))
        dw_s, dw_v = be.node_gemm_wgrad(x0, x1, dy, None, 0)
        return dict(w_s=dw_s, w_v=dw_v, bias=dbias)
    grads = side.run(weight_side) if side is not None else weight_side()
    k = w["w_s"].shape[0]
    dx = be.node_gemm(dy, None, _transposed(w), k, split=split_out or 0)
    return dx, grads


def backward_train(W: Dict, saved: Dict, dpred, backend=None):
    """Returns the gradient tree matching W (None where a leaf has no gradient, e.g. BatchNorm buffers)."""
    be = backend or _ops
    colsum2 = _colsum2_of(be)
    n, B, N = saved["n"], saved["B"], saved["N"]
    nodes, deg = B * N, N - 1
    E = nodes * deg
    attr, pos, mass = saved["attr"], saved["pos"], saved["mass"]
    bn_training = saved["bn_training"]
    grads = dict(layers=[])
    side = _SideWork(dpred) if backend is None else None  # test back ends (CPU emulation) run in program order
    # head + pre_pool1
    dhp, dw_head = be.head_bwd(saved["hp"], attr, W["head"], dpred.contiguous(), n)
    grads["head"] = dw_head
    dh, grads["pool1"] = _node_tp_backward(be, W["pool1"], saved["h_last"], None, saved["yp"], attr, n, True, dhp,
                                           side=side)
    for lw, rec in zip(reversed(W["layers"]), reversed(saved["layers"])):
        g = dict(bn_msg=None, bn_feat=None)
        m1, m2, u1, u2 = lw["msg1"], lw["msg2"], lw["upd1"], lw["upd2"]
        # feature BatchNorm
        if lw["bn_feat"] is not None:
            flat_pre, flat_dh = rec["pre"].view(nodes, 4 * n), dh.view(nodes, 4 * n)
            sum_g, sum_gx = colsum2(flat_dh, None, 0, flat_dh, flat_pre, 2)
            c = be.bn_backward_coeffs(lw["bn_feat"], rec["bn_feat"], n, float(nodes), 1.0, sum_g, sum_gx, bn_training)
            dpre = be.lincomb(flat_dh, flat_pre, c["A4"], c["B4"], c["C4"]).view(nodes, 4, n)
            g["bn_feat"] = dict(weight=c["dweight"], bias=c["dbias"])
        else:
            dpre = dh
        # update_layer_2 (+ residual), update_layer_1
        dg1, g["upd2"] = _node_tp_backward(be, u2, rec["g1"], None, rec["y2"], attr, n, False, dpre, side=side)
        (dh_u, dagg), g["upd1"] = _node_tp_backward(be, u1, rec["h"], rec["agg"], rec["y1"], attr, n, True, dg1,
                                                    split_out=n, side=side)
        # message BatchNorm folded through the sum over senders
        if lw["bn_msg"] is not None:
            flat_g, flat_raw = dagg.view(nodes, 4 * n), rec["agg_raw"].view(nodes, 4 * n)
            sum_g, sum_gx = colsum2(flat_g, None, 0, flat_g, flat_raw, 2)
            c = be.bn_backward_coeffs(lw["bn_msg"], rec["bn_msg"], n, float(E), float(deg), sum_g, sum_gx, bn_training)
            g["bn_msg"] = dict(weight=c["dweight"], bias=c["dbias"])
            bn_a, bn_b, bn_c = c["bn_a"], c["bn_b"], c["bn_c"]
        else:
            bn_a = torch.ones(2 * n, dtype=dagg.dtype, device=dagg.device)
            bn_b, bn_c = torch.zeros_like(bn_a), torch.zeros(n, dtype=dagg.dtype, device=dagg.device)
        # fused edge layer backward (recompute) -> dP, dQ, message_layer_2 and w_edge gradients
        if rec.get("rows") is not None:  # the forward call left this layer's edge rows in HBM: no recompute
            dP, dQ, g["msg2"], dwe = _ops.edge_layer_gemm_bwd(pos, mass, B, N, n, rec["p"], rec["q"], m1["w_edge"], m2,
                                                              bn_a, bn_b, bn_c, dagg, rows=rec["rows"], side=side)
            rec["rows"] = None  # 11 n floats per edge row go back to the allocator as soon as the layer is done
        else:
            kw = dict(side=side) if side is not None else {}  # test back ends keep the plain signature
            dP, dQ, g["msg2"], dwe = be.edge_layer_bwd(pos, mass, B, N, n, rec["p"], rec["q"], m1["w_edge"], m2, bn_a,
                                                       bn_b, bn_c, dagg, **kw)
        # message_layer_1 projections: [P | Q] = h @ W (+ bias on P's l=0 columns)
        def msg1_weights(h_=rec["h"], dP_=dP, dQ_=dQ):
            dw_s, dw_v = be.node_gemm_wgrad(h_, None, dP_, dQ_, 3 * n)
            dbias1 = be.colsum(dP_.view(nodes, 12 * n))[:2 * n]
            return dict(w_s=dw_s, w_v=dw_v, bias=dbias1.contiguous())
        g["msg1"] = side.run(msg1_weights) if side is not None else msg1_weights()
        g["msg1"]["w_edge"] = dwe
        dh_m = be.node_gemm(dP, dQ, _transposed(m1), n)
        dh = be.add3(dpre, dh_u, dh_m)
        grads["layers"].append(g)
    grads["layers"].reverse()
    dw_e, db_e = be.embed_bwd(saved["x_in"], attr, dh, n)
    grads["embed"] = dict(w=dw_e, bias=db_e)
    if side is not None:
        side.join()
    return grads


def attach_bn_buffers(W, bufs):
    """BatchNorm running statistics / eps / momentum ride next to the differentiable leaves."""
    for lw, lb in zip(W["layers"], bufs):
        for key in ("bn_msg", "bn_feat"):
            if lw[key] is not None:
                lw[key].update(lb[key])


# ---------------------------------------------------------------------------------------------------------------
# parameters -> packed operand blocks as ONE gather
# ---------------------------------------------------------------------------------------------------------------
def build_pack_map(model):
    """The re-layout of the reference-named parameters into the kernels' operand blocks (packing.py) is a fixed linear
    map in which every packed element is one parameter element times a constant. Run through autograd it costs ~1500
    tiny launches per training step (a zero-fill, a copy and an add per view in the backward): more than the arithmetic
    of the README configuration. This derives the map once as gather tables by probing ``model.packed_train`` in
    float64 with two parameter settings (all ones -> the constants; distinct integer ids -> the source elements), so
    that packing is ``cat(params)[IDX] * SCALE`` and its transpose one ``index_add_``.
    Returns dict(idx, scale, spec, shapes, sizes, param_sizes, ...); BatchNorm buffers are NOT part of the map."""
    params = list(model.parameters())
    sizes = [p.numel() for p in params]
    total = sum(sizes)
    if total >= 2 ** 24:
        raise NotImplementedError("pack map probing needs fewer than 2^24 parameters (fp32-exact integer ids)")
    with torch.no_grad():
        saved = [p.detach().clone() for p in params]
        try:
            for p in params:
                p.fill_(1.0)
            ones_leaves, spec = flatten_packed(model.packed_train(torch.float64, transposed=True)[0])
            ones_leaves = [t.clone() for t in ones_leaves]  # pass-through leaves alias the parameters themselves
            off = 0
            for p, n in zip(params, sizes):
                p.copy_(torch.arange(off + 1, off + n + 1, device=p.device, dtype=torch.float64).reshape(p.shape))
                off += n
            id_leaves, _ = flatten_packed(model.packed_train(torch.float64, transposed=True)[0])
            id_leaves = [t.clone() for t in id_leaves]
        finally:
            for p, v in zip(params, saved):
                p.copy_(v)
        scale = torch.cat([t.reshape(-1) for t in ones_leaves])
        ids = torch.cat([t.reshape(-1) for t in id_leaves])
        safe = torch.where(scale != 0, scale, torch.ones_like(scale))
        idx = torch.where(scale != 0, torch.round(ids / safe) - 1, torch.zeros_like(ids)).to(torch.int64)
        if int(idx.min()) < 0 or int(idx.max()) >= total:
            raise AssertionError("pack map probe produced an out-of-range source index")
        # every packed element must be exactly scale * parameter[idx]
        if float((ids - scale * (idx + 1).to(torch.float64)).abs().max()) > 1e-6:
            raise AssertionError("packing is not a gather with constants: pack map unusable")
    # leaves that can carry a gradient (the transposed copies "*_t" are derived: the backward kernels only read them)
    derived = set()

    def mark(node):
        if "d" in node:
            for k, sub in node["d"]:
                if k.endswith("_t") and "t" in sub:
                    derived.add(sub["t"])
                else:
                    mark(sub)
        elif "l" in node:
            for sub in node["l"]:
                mark(sub)

    mark(spec)
    leaf_sizes = [t.numel() for t in ones_leaves]
    offsets = [0]
    for nleaf in leaf_sizes:
        offsets.append(offsets[-1] + nleaf)
    grad_leaves = [i for i in range(len(leaf_sizes)) if i not in derived]
    sel = torch.cat([torch.arange(offsets[i], offsets[i + 1], device=idx.device) for i in grad_leaves])
    return dict(idx=idx, scale=scale, spec=spec, shapes=[tuple(t.shape) for t in ones_leaves], sizes=leaf_sizes,
                param_sizes=sizes, grad_leaves=grad_leaves, grad_idx=idx[sel].contiguous(),
                grad_scale=scale[sel].contiguous())


class SegnnTrainFunctionFlat(torch.autograd.Function):
    """pred = SEGNN(pos, vel, mass; parameters) with the hand-written backward, taking the reference-named parameters
    themselves: packing is one gather (build_pack_map) inside the function, its transpose one index_add_ in backward."""

    @staticmethod
    def forward(ctx, cfg, pos, vel, mass, *params):
        pm, dtype = cfg["pack_map"], cfg["dtype"]
        flat = cfg.get("flat_params")  # SEGNN.use_flat_storage: the parameters ARE slices of this buffer
        if flat is None:
            flat = torch.cat([p.detach().reshape(-1) for p in params]).to(dtype)
        scale = pm["scale"].to(dtype)
        packed = flat.index_select(0, pm["idx"]) * scale
        leaves = [t.view(shape) for t, shape in zip(packed.split(pm["sizes"]), pm["shapes"])]
        W = unflatten_packed(leaves, pm["spec"])
        # the buffers are read from the module on every call (cfg), never from the cached map: .float()/.double()/
        # load_state_dict(assign=True) replace the buffer objects without changing the map's key
        attach_bn_buffers(W, cfg["bn_buffers"])
        pred, saved = forward_train(W, cfg["n"], pos, vel, mass, cfg["B"], cfg["N"], cfg["bn_training"],
                                    backend=cfg.get("backend"), keep_rows=True)
        ctx.cfg, ctx.W, ctx.saved, ctx.scale = cfg, W, saved, scale
        ctx.param_shapes = [tuple(p.shape) for p in params]
        ctx.param_dtypes = [p.dtype for p in params]
        return pred

    @staticmethod
    def backward(ctx, dpred):
        pm = ctx.cfg["pack_map"]
        grads = backward_train(ctx.W, ctx.saved, dpred, backend=ctx.cfg.get("backend"))
        gl, gspec = flatten_packed(_align(grads, pm["spec"]))
        out = [None] * len(pm["sizes"])
        _scatter_leaves(gspec, pm["spec"], gl, out)
        dev, dt = ctx.scale.device, ctx.scale.dtype
        pieces = [out[i].reshape(-1).to(dt) if out[i] is not None
                  else torch.zeros(pm["sizes"][i], dtype=dt, device=dev) for i in pm["grad_leaves"]]
        sink = ctx.cfg.get("grad_sink")
        if sink is not None:  # every .grad is a slice of `sink`: one scatter-add writes all of them, autograd gets None
            sink.zero_()
            sink.index_add_(0, pm["grad_idx"], torch.cat(pieces) * pm["grad_scale"].to(dt))
            ctx.saved = None
            return (None, None, None, None) + (None,) * len(ctx.param_shapes)
        gflat = torch.zeros(sum(pm["param_sizes"]), dtype=dt, device=dev)
        gflat.index_add_(0, pm["grad_idx"], torch.cat(pieces) * pm["grad_scale"].to(dt))
        gparams = [g.view(shape).to(pdt) for g, shape, pdt in
                   zip(gflat.split(pm["param_sizes"]), ctx.param_shapes, ctx.param_dtypes)]
        ctx.saved = None
        return (None, None, None, None, *gparams)


class SegnnTrainFunction(torch.autograd.Function):
    """pred = SEGNN(pos, vel, mass; packed weights) with the hand-written backward. The packed weights are
    differentiable torch re-layouts of the reference-named parameters (packing.py), so autograd carries the packed
    gradients back to ``tp.weight`` / ``biases`` / BatchNorm ``weight`` / ``bias``."""

    @staticmethod
    def forward(ctx, cfg, pos, vel, mass, *leaves):
        W = unflatten_packed([t.detach() for t in leaves], cfg["spec"])
        attach_bn_buffers(W, cfg["bn_buffers"])
        pred, saved = forward_train(W, cfg["n"], pos, vel, mass, cfg["B"], cfg["N"], cfg["bn_training"],
                                    backend=cfg.get("backend"), keep_rows=True)
        ctx.cfg, ctx.W, ctx.saved, ctx.n_leaves = cfg, W, saved, len(leaves)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        grads = backward_train(ctx.W, ctx.saved, dpred, backend=ctx.cfg.get("backend"))
        gl, gspec = flatten_packed(_align(grads, ctx.cfg["spec"]))
        out = [None] * ctx.n_leaves
        _scatter_leaves(gspec, ctx.cfg["spec"], gl, out)
        ctx.saved = None
        return (None, None, None, None, *out)


def _align(grads, spec):
    """Shape the gradient tree like the weight tree (missing entries -> None)."""
    if "d" in spec:
        return {k: _align(grads.get(k) if isinstance(grads, dict) else None, s) for k, s in spec["d"]}
    if "l" in spec:
        return [_align(grads[i] if grads is not None else None, s) for i, s in enumerate(spec["l"])]
    return grads


def _scatter_leaves(gspec, wspec, gleaves, out):
    if "d" in wspec:
        gd = dict(gspec["d"]) if "d" in gspec else {}
        for k, s in wspec["d"]:
            if k in gd:
                _scatter_leaves(gd[k], s, gleaves, out)
    elif "l" in wspec:
        if "l" in gspec:
            for gs, s in zip(gspec["l"], wspec["l"]):
                _scatter_leaves(gs, s, gleaves, out)
    elif "t" in wspec and "t" in gspec:
        out[wspec["t"]] = gleaves[gspec["t"]]
