"""Generic-irreps forward of the SEGNN modules (fp32, inference): the reference's own formulation -- gathered message
inputs, one tensor product per call, e3nn Gate, BatchNorm (eval), scatter-sum -- on plain CUDA kernels
(csrc/segnn_generic.cu) that accept any hidden irreps with l <= 2 and lmax_attr <= 2. It gives parity coverage for configurations
the fused kernels are not specialised for (lmax_h = 2, BASELINE config 3); per-edge tensors live in HBM here, as in the
reference. Per-edge tensor products run as expansion kernel + plain fp32 library GEMM + scatter kernel (the weight
contraction is a plain GEMM once the coupling with the edge attribute is applied); node-level ones as one kernel.

models/segnn/segnn.py:150-189 (SEGNN.forward), :239-304 (SEGNNLayer), o3_building_blocks.py:150-203.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict

import numpy as np
import torch

from . import ops
from ._lib import check, lib
from .cg import real_wigner_3j
from .irreps import Irreps

_p = ops._p


def _slices(irreps: Irreps):
    out, off = [], 0
    for m, l, _ in irreps:
        out.append((off, m, 2 * l + 1))
        off += m * (2 * l + 1)
    return out


class TensorProductPlan:
    """Device-side description of one O3TensorProduct: instruction table, scaled couplings, dense bias."""

    def __init__(self, module, device):
        tp = module.tp
        s1, s2, so = _slices(tp.irreps_in1), _slices(tp.irreps_in2), _slices(tp.irreps_out)
        instr, cgs = [], []
        for ins in tp.instructions:
            l1, l2, lo = ins["ls"]
            if max(l1, l2, lo) > 2 or ins["shape"][1] != 1:
                raise NotImplementedError("generic tensor product: l <= 2 for features and attributes, attribute "
                                          "multiplicity 1")
            o1, m1, d1 = s1[ins["i1"]]
            o2, _, d2 = s2[ins["i2"]]
            oo, mo, do = so[ins["io"]]
            instr.append([o1, m1, d1, o2, d2, oo, mo, do, ins["offset"]])
            c = np.zeros((5, 5, 5), dtype=np.float64)
            # net path coefficient: e3nn 'component'/'element' path weight x the reference's sqrt_k_correction
            c[:d1, :d2, :do] = math.sqrt(2 * lo + 1) * real_wigner_3j(l1, l2, lo)
            cgs.append(c)
        # "expand + GEMM" form (large row counts): per output irrep block the paths that write it, with the column
        # offset of each path inside the stacked weight matrix [K][mulo]
        self.blocks = []
        for io, (oo, mo, do) in enumerate(so):
            paths, cg_io, wviews, koff = [], [], [], 0
            for q, ins in enumerate(tp.instructions):
                if ins["io"] != io:
                    continue
                o1, m1, d1 = s1[ins["i1"]]
                o2, _, d2 = s2[ins["i2"]]
                paths.append([o1, m1, d1, o2, d2, koff])
                cg_io.append(cgs[q])
                wviews.append((ins["offset"], m1))
                koff += m1
            if paths:
                self.blocks.append(dict(
                    paths=torch.tensor(paths, dtype=torch.int32, device=device).contiguous(),
                    cg=torch.tensor(np.stack(cg_io), dtype=torch.float32, device=device).contiguous(),
                    n_paths=len(paths), wviews=wviews, K=koff, offo=oo, mulo=mo, dimo=do))
        self.n_instr = len(instr)
        self.instr = torch.tensor(instr, dtype=torch.int32, device=device).contiguous()
        self.cg = torch.tensor(np.stack(cgs), dtype=torch.float32, device=device).contiguous()
        self.d1, self.d2, self.dout = tp.irreps_in1.dim, tp.irreps_in2.dim, tp.irreps_out.dim
        cols = [c for (off, m, d), (_, l, _) in zip(so, tp.irreps_out) if l == 0 for c in range(off, off + m)]
        self.bias_idx = torch.tensor(cols, dtype=torch.int64, device=device) if cols else None
        self.module = module

    def run(self, x1, x2):
        rows = x1.shape[0]
        m = self.module
        w = m.tp.weight.detach().to(torch.float32).contiguous()
        bias = None
        if m.biases is not None:
            bias = torch.zeros(self.dout, dtype=torch.float32, device=x1.device)
            bias[self.bias_idx] = m.biases.detach().to(torch.float32)
        out = torch.empty((rows, self.dout), dtype=torch.float32, device=x1.device)
        if rows >= self.GEMM_MIN_ROWS:
            return self._run_gemm(x1, x2, w, bias, out)
        with torch.cuda.device(x1.device):
            check(lib.segnn_generic_tp(_p(x1), self.d1, _p(x2), self.d2, rows, _p(w), _p(self.instr), self.n_instr,
                                       _p(self.cg), _p(bias), self.dout, _p(out), ops._stream()), "segnn_generic_tp")
        ops._bump()
        return out

    GEMM_MIN_ROWS = 128       # below this the single tiled kernel is used (too few rows to fill a GEMM)
    GEMM_CHUNK_BYTES = 2 << 30  # bound on the expanded operand A per chunk of rows

    def _run_gemm(self, x1, x2, w, bias, out):
        """Per output irrep block: expansion kernel (coupling with x2 applied) -> fp32-accurate tcgen05 GEMM
        (segnn_gemm_tf32x3) against the stacked path weights -> scatter kernel into the e3nn column order. Same
        arithmetic as segnn_generic_tp, with the weight contraction on the tensor cores instead of one thread per
        output."""
        rows = x1.shape[0]
        if sum(b["mulo"] * b["dimo"] for b in self.blocks) != self.dout:
            out.zero_()  # output columns no path writes (none for the reference's products) are bias / 0
            if bias is not None:
                out += bias
        for b in self.blocks:
            wcat = torch.cat([w[off: off + m1 * b["mulo"]].view(m1, b["mulo"]) for off, m1 in b["wviews"]], dim=0)
            per_row = b["dimo"] * b["K"] * 4
            chunk = max(1, min(rows, self.GEMM_CHUNK_BYTES // per_row))
            for r0 in range(0, rows, chunk):
                r1 = min(rows, r0 + chunk)
                n = r1 - r0
                # rows padded to a multiple of four floats: 128-bit accesses in the GEMM (padding is never read as data:
                # the loader masks k >= K, the scatter reads mulo columns)
                lda, ldy = (b["K"] + 3) & ~3, (b["mulo"] + 3) & ~3
                A = torch.empty((n * b["dimo"], lda), dtype=torch.float32, device=x1.device)
                Y = torch.empty((n * b["dimo"], ldy), dtype=torch.float32, device=x1.device)
                with torch.cuda.device(x1.device):
                    check(lib.segnn_generic_tp_expand_ld(_p(x1[r0:r1]), self.d1, _p(x2[r0:r1]), self.d2, n,
                                                         _p(b["paths"]), b["n_paths"], _p(b["cg"]), b["dimo"], b["K"],
                                                         lda, _p(A), ops._stream()), "segnn_generic_tp_expand_ld")
                    # the weight contraction: own tcgen05 GEMM, fp32-accurate (3xTF32), no library call
                    ops.gemm_tf32x3(A[:, :b["K"]], wcat, out=Y[:, :b["mulo"]])
                    check(lib.segnn_generic_tp_scatter_ld(_p(Y), ldy, n, b["dimo"], b["mulo"], b["offo"], self.dout,
                                                          _p(bias), _p(out[r0:r1]), ops._stream()),
                          "segnn_generic_tp_scatter_ld")
                ops._bump(2)
        return out


class HoistedMessage1Plan:
    """message_layer_1 (segnn.py:212-214, 264-279) with the weight contraction of its x_i / x_j parts hoisted to node
    level: the tensor product is linear in cat(x_i, x_j, add), so Y = W^T x is computed once per node (generic tensor
    product with an identity coupling) and per edge only the coupling with the edge attribute remains
    (segnn_generic_hoisted_msg1). Same result as TensorProductPlan(message_layer_1) on the gathered message input,
    without the [E, 2D + 2] input tensor and with ~25x fewer multiply-adds per edge."""

    def __init__(self, module, hidden_irreps: Irreps, device):
        tp = module.tp
        nh, D = len(hidden_irreps), hidden_irreps.dim
        s1, s2, so = _slices(tp.irreps_in1), _slices(tp.irreps_in2), _slices(tp.irreps_out)
        yinstr, ycg, slots, adds, cgs_add, yoff = [], [], {}, [], [], 0
        for ins in tp.instructions:
            l1, l2, lo = ins["ls"]
            if max(l1, l2, lo) > 2 or ins["shape"][1] != 1:
                raise NotImplementedError("generic tensor product: l <= 2 for features and attributes")
            o1, m1, d1 = s1[ins["i1"]]
            o2, _, d2 = s2[ins["i2"]]
            oo, mo, do = so[ins["io"]]
            c = np.zeros((5, 5, 5), dtype=np.float64)
            c[:d1, :d2, :do] = math.sqrt(2 * lo + 1) * real_wigner_3j(l1, l2, lo)
            if ins["i1"] >= 2 * nh:  # additional_message_features block (scalars)
                if l1 != 0:
                    raise NotImplementedError("additional message features must be scalars")
                adds.append([oo, mo, do, o2, d2, ins["offset"], m1])
                cgs_add.append(c)
                continue
            role = ins["i1"] // nh  # 0: x_i (receiver), 1: x_j (sender)
            ident = np.zeros((5, 5, 5), dtype=np.float64)
            for i in range(d1):
                ident[i, 0, i] = 1.0
            yinstr.append([o1 - role * D, m1, d1, 0, 1, yoff, mo, d1, ins["offset"]])
            ycg.append(ident)
            key = (ins["i1"] % nh, ins["i2"], ins["io"])
            slot = slots.setdefault(key, dict(meta=[oo, mo, do, d1, o2, d2], cg=c, yoff=[None, None]))
            slot["yoff"][role] = yoff
            yoff += mo * d1
        pairs, cgs = [], []
        for slot in slots.values():
            if None in slot["yoff"]:
                raise NotImplementedError("x_i / x_j instructions of message_layer_1 do not pair up")
            pairs.append(slot["meta"] + slot["yoff"])
            cgs.append(slot["cg"])
        t32 = lambda a: torch.tensor(a, dtype=torch.int32, device=device).contiguous()
        self.ydim, self.n_y = yoff, len(yinstr)
        self.yinstr = t32(yinstr)
        self.ycg = torch.tensor(np.stack(ycg), dtype=torch.float32, device=device).contiguous()
        self.n_pairs, self.n_adds = len(pairs), len(adds)
        self.pairs = t32(pairs)
        self.adds = t32(adds) if adds else None
        self.cg = torch.tensor(np.stack(cgs + cgs_add), dtype=torch.float32, device=device).contiguous()
        self.blocks = t32([[oo, mo, do] for oo, mo, do in so])
        self.n_blocks, self.n_items = len(so), sum(mo for _, mo, _ in so)
        self.D, self.d2, self.dout = D, tp.irreps_in2.dim, tp.irreps_out.dim
        cols = [c for (off, m, d), (_, l, _) in zip(so, tp.irreps_out) if l == 0 for c in range(off, off + m)]
        self.bias_idx = torch.tensor(cols, dtype=torch.int64, device=device) if cols else None
        self.module = module

    def run(self, x, edge_attr, add, B: int, N: int):
        m, dev = self.module, x.device
        nodes = x.shape[0]
        w = m.tp.weight.detach().to(torch.float32).contiguous()
        bias = None
        if m.biases is not None:
            bias = torch.zeros(self.dout, dtype=torch.float32, device=dev)
            bias[self.bias_idx] = m.biases.detach().to(torch.float32)
        ones = torch.ones((nodes, 1), dtype=torch.float32, device=dev)
        Y = torch.empty((nodes, self.ydim), dtype=torch.float32, device=dev)
        out = torch.empty((edge_attr.shape[0], self.dout), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.segnn_generic_tp(_p(x), self.D, _p(ones), 1, nodes, _p(w), _p(self.yinstr), self.n_y, _p(self.ycg),
                                       None, self.ydim, _p(Y), ops._stream()), "segnn_generic_tp")
            check(lib.segnn_generic_hoisted_msg1(_p(Y), self.ydim, _p(edge_attr), self.d2, _p(add), add.shape[1], B, N,
                                                 _p(self.pairs), self.n_pairs, _p(self.adds), self.n_adds,
                                                 _p(self.blocks), self.n_blocks, self.n_items, _p(self.cg), _p(w),
                                                 _p(bias), self.dout, _p(out), ops._stream()),
                  "segnn_generic_hoisted_msg1")
        ops._bump(2)
        return out


class GatePlan:
    """e3nn Gate layout of an O3TensorProductSwishGate output (o3_building_blocks.py:175-193)."""

    def __init__(self, module, device):
        out = module.irreps_gated_out
        self.n_s = out[0][0] if out[0][1] == 0 else 0
        gated = out[1:] if self.n_s else out
        idx, g = [], 0
        for m, l, _ in gated:
            for u in range(m):
                idx += [g] * (2 * l + 1)
                g += 1
        self.n_g, self.d_gated = g, len(idx)
        self.gate_index = torch.tensor(idx, dtype=torch.int32, device=device) if idx else None

    def run(self, x):
        rows = x.shape[0]
        out = torch.empty((rows, self.n_s + self.d_gated), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.segnn_generic_gate(_p(x), rows, self.n_s, self.n_g, self.d_gated, _p(self.gate_index), _p(out),
                                         ops._stream()), "segnn_generic_gate")
        ops._bump()
        return out


def _bn_eval_columns(bn, irreps: Irreps):
    """Eval-mode e3nn BatchNorm (segnn.py:233-235) as a per-column affine (mul, add)."""
    f = lambda t: t.detach().to(torch.float32)
    mul_irrep = f(bn.weight) * (f(bn.running_var) + bn.eps).rsqrt()
    mul, add, ii, isc = [], [], 0, 0
    for m, l, p in irreps:
        d = 2 * l + 1
        mi = mul_irrep[ii: ii + m]
        mul.append(mi.repeat_interleave(d))
        if l == 0 and p == 1:
            add.append(f(bn.bias)[isc: isc + m] - f(bn.running_mean)[isc: isc + m] * mi)
            isc += m
        else:
            add.append(torch.zeros(m * d, dtype=torch.float32, device=mi.device))
        ii += m
    return torch.cat(mul).contiguous(), torch.cat(add).contiguous()


def _feature_norm(layer, x, B: int, N: int):
    """segnn.py:257-261: eval-mode BatchNorm (a per-column affine) or InstanceNorm over the N nodes of each graph."""
    if layer.norm == "instance":
        return layer.feature_norm.forward_equal_graphs(x, B, N)
    mul, addc = _bn_eval_columns(layer.feature_norm, layer.hidden_irreps)
    return ops.lincomb(x, None, mul, None, addc)


USE_L2_ROWS = True  # False: every layer through the table-driven generic kernels (the checker of the l2 form)
_L2_TYPES = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 2), (2, 0, 2), (2, 1, 1)]


class Lmax2EdgePlan:
    """Edge part of one SEGNNLayer for hidden irreps n x 0e + n x 1o + n x 2e (lmax_h = 2, BASELINE configuration 3) in
    GEMM form (csrc/segnn_l2_rows.cu): node-level products of message_layer_1 (the hoisted plan's Y), ONE kernel for
    message_layer_1's per-edge coupling + gate + message_layer_2's coupling -> three GEMM operands, the package's
    fp32-accurate tcgen05 GEMM against the stacked path weights, ONE kernel for bias + gate + sum over senders + eval
    BatchNorm.  Same arithmetic as the generic kernels (which remain the checker: tests/test_gpu_parity.py), without
    the [E, .] tensors between hoisted message_layer_1, gate, expansion, scatter, gate, BatchNorm and aggregation."""

    CHUNK_BYTES = 4 << 30

    @staticmethod
    def supported(layer, hidden_irreps: Irreps) -> bool:
        try:
            h = [(m, l, p) for m, l, p in hidden_irreps]
            if len(h) != 3 or [l for _, l, _ in h] != [0, 1, 2] or len({m for m, _, _ in h}) != 1 or h[0][0] > 96:
                return False
            if [p for _, _, p in h] != [1, -1, 1]:
                return False
            t1, t2 = layer.message_layer_1.tp, layer.message_layer_2.tp
            attr = [(m, l) for m, l, _ in t1.irreps_in2]
            if attr != [(1, 0), (1, 1)] or [(m, l) for m, l, _ in t2.irreps_in2] != attr:
                return False
            n = h[0][0]
            outs = [(m, l) for m, l, _ in t1.irreps_out]
            if outs != [(3 * n, 0), (n, 1), (n, 2)] or [(m, l) for m, l, _ in t2.irreps_out] != outs:
                return False
            in1 = [(m, l) for m, l, _ in t1.irreps_in1]
            return in1 == [(n, 0), (n, 1), (n, 2)] * 2 + [(2, 0)] and \
                [(m, l) for m, l, _ in t2.irreps_in1] == [(n, 0), (n, 1), (n, 2)]
        except Exception:
            return False

    def __init__(self, layer, hidden_irreps: Irreps, hoisted: "HoistedMessage1Plan", msg2: TensorProductPlan, device):
        self.layer, self.hoisted, self.msg2 = layer, hoisted, msg2
        self.n = n = hidden_irreps[0][0]
        tindex = {t: i for i, t in enumerate(_L2_TYPES)}
        cg = np.zeros((7, 5, 3, 5), dtype=np.float64)
        for (l1, l2, lo), i in tindex.items():
            cg[i, :2 * l1 + 1, :2 * l2 + 1, :2 * lo + 1] = math.sqrt(2 * lo + 1) * real_wigner_3j(l1, l2, lo)
        self.cg = torch.tensor(cg, dtype=torch.float32, device=device).contiguous()
        ltype = lambda d1, d2, do: tindex[((d1 - 1) // 2, (d2 - 1) // 2, (do - 1) // 2)]
        # node-level products of message_layer_1 as three GEMMs (one per input degree l1): the weight views [n][mo] of
        # the instructions with that input degree, x_i and x_j roles side by side; yoff[type][role] = first column
        nh = len(hidden_irreps)
        self.ystack = {0: [], 1: [], 2: []}  # l1 -> [(weight offset, mo)]
        ncols = {0: 0, 1: 0, 2: 0}
        yoff = [[-1, -1] for _ in range(7)]
        for ins in layer.message_layer_1.tp.instructions:
            if ins["i1"] >= 2 * nh:
                continue
            role, (l1, l2, lo) = ins["i1"] // nh, ins["ls"]
            mo = ins["shape"][2]
            assert ins["shape"][0] == n
            yoff[tindex[(l1, l2, lo)]][role] = ncols[l1]
            self.ystack[l1].append((ins["offset"], mo))
            ncols[l1] += mo
        assert all(a >= 0 and b >= 0 for a, b in yoff)
        self.ncols = ncols
        self.yoff = torch.tensor(yoff, dtype=torch.int32).contiguous()  # host: read by the C call, not by a kernel
        self.add_off = {}
        for oo, mo, do, o2, d2, woff, mul1 in hoisted.adds.cpu().tolist():
            assert mul1 == 2
            self.add_off[ltype(1, d2, do)] = (woff, mo)
        assert set(self.add_off) == {0, 1} and self.add_off[0][1] == 3 * n and self.add_off[1][1] == n
        koff = [-1] * 7
        self.blocks = {}
        for b in msg2.blocks:
            self.blocks[b["dimo"]] = b
            for o1, m1, d1, o2, d2, ko in b["paths"].cpu().tolist():
                assert m1 == n
                koff[ltype(d1, d2, b["dimo"])] = ko
        assert all(k >= 0 for k in koff) and set(self.blocks) == {1, 3, 5}
        assert self.blocks[1]["mulo"] == 3 * n and self.blocks[3]["mulo"] == n and self.blocks[5]["mulo"] == n
        self.koff = torch.tensor(koff, dtype=torch.int32).contiguous()
        self._cache_key, self._cache = None, None

    def _weights(self, device):
        m1, m2 = self.layer.message_layer_1, self.layer.message_layer_2
        norm = self.layer.message_norm
        key = (m1.tp.weight._version, m2.tp.weight._version, m1.biases._version, m2.biases._version,
               None if norm is None else (norm.weight._version, norm.bias._version, norm.running_mean._version,
                                          norm.running_var._version))
        if key == self._cache_key:
            return self._cache
        n = self.n
        w1 = m1.tp.weight.detach().to(torch.float32).contiguous()
        w2 = m2.tp.weight.detach().to(torch.float32).contiguous()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        c = dict(tp_weight_1=w1, w_add0=w1[self.add_off[0][0]: self.add_off[0][0] + 2 * 3 * n],
                 w_add1=w1[self.add_off[1][0]: self.add_off[1][0] + 2 * n],
                 bias1=f32(m1.biases)[: 3 * n], bias2=f32(m2.biases)[: 3 * n])
        assert m1.biases.numel() == 3 * n and m2.biases.numel() == 3 * n
        for l1, views in self.ystack.items():
            c[f"ystack{l1}"] = torch.cat([w1[off: off + n * mo].view(n, mo) for off, mo in views], dim=1).contiguous()
        for d, b in self.blocks.items():
            c[f"stacked{d}"] = torch.cat([w2[off: off + k * b["mulo"]].view(k, b["mulo"]) for off, k in b["wviews"]],
                                   dim=0).contiguous()
        c["bn"] = _bn_eval_columns(norm, self.layer.hidden_irreps) if norm is not None else (None, None)
        self._cache_key, self._cache = key, c
        return c

    def run(self, x, pos, mass, B: int, N: int):
        dev, n = x.device, self.n
        c = self._weights(dev)
        hp = self.hoisted
        nodes = B * N
        agg = torch.empty((nodes, 9 * n), dtype=torch.float32, device=dev)
        K = {d: self.blocks[d]["K"] for d in (1, 3, 5)}
        lda = {d: (K[d] + 3) & ~3 for d in (1, 3, 5)}
        ldy = {1: (3 * n + 3) & ~3, 3: (n + 3) & ~3, 5: (n + 3) & ~3}
        per_graph = 4 * N * N * sum(d * (lda[d] + ldy[d]) for d in (1, 3, 5))
        gpc = max(1, min(B, self.CHUNK_BYTES // per_graph))
        ldx = (n + 3) & ~3
        xp = [torch.empty((nodes * d, ldx), dtype=torch.float32, device=dev) for d in (1, 3, 5)]
        ldn = [(self.ncols[l1] + 3) & ~3 for l1 in (0, 1, 2)]
        Yn = [torch.empty((nodes * d, ldn[l1]), dtype=torch.float32, device=dev) for l1, d in enumerate((1, 3, 5))]
        with torch.cuda.device(dev):
            check(lib.segnn_l2_planarize(_p(x), nodes, n, ldx, _p(xp[0]), _p(xp[1]), _p(xp[2]), ops._stream()),
                  "segnn_l2_planarize")
            ops._bump()
            for l1 in (0, 1, 2):
                ops.gemm_tf32x3(xp[l1][:, :n], c[f"ystack{l1}"], out=Yn[l1][:, :self.ncols[l1]])
            for g0 in range(0, B, gpc):
                gc = min(gpc, B - g0)
                rows = gc * N * N
                A = {d: torch.empty((rows * d, lda[d]), dtype=torch.float32, device=dev) for d in (1, 3, 5)}
                Yb = {d: torch.empty((rows * d, ldy[d]), dtype=torch.float32, device=dev) for d in (1, 3, 5)}
                check(lib.segnn_l2_msg_rows(_p(pos), _p(mass), gc, N, n, g0 * N, _p(Yn[0]), ldn[0], _p(Yn[1]), ldn[1],
                                            _p(Yn[2]), ldn[2], self.yoff.data_ptr(), _p(self.cg), _p(c["w_add0"]), _p(c["w_add1"]),
                                            _p(c["bias1"]), self.koff.data_ptr(), lda[1], lda[3], lda[5], _p(A[1]),
                                            _p(A[3]), _p(A[5]), ops._stream()), "segnn_l2_msg_rows")
                for d in (1, 3, 5):
                    ops.gemm_tf32x3(A[d][:, :K[d]], c[f"stacked{d}"], out=Yb[d][:, :self.blocks[d]["mulo"]])
                mul, addc = c["bn"]
                check(lib.segnn_l2_gate_aggregate(gc, N, n, g0 * N, _p(Yb[1]), ldy[1], _p(Yb[3]), ldy[3], _p(Yb[5]),
                                                  ldy[5], _p(c["bias2"]), _p(mul), _p(addc), _p(agg), ops._stream()),
                      "segnn_l2_gate_aggregate")
                ops._bump(2)
        return agg


class GenericRunner:
    """Plans for every tensor product / gate of one SEGNN; ``forward`` mirrors SEGNN.forward (eval mode)."""

    def __init__(self, model, device):
        self.model = model
        tp = lambda m: TensorProductPlan(m, device)
        self.embed = tp(model.embedding_layer)
        self.layers = []
        for layer in model.layers:
            self.layers.append(dict(
                msg1=tp(layer.message_layer_1), g_msg1=GatePlan(layer.message_layer_1, device),
                msg1h=HoistedMessage1Plan(layer.message_layer_1, model.hidden_irreps, device),
                msg2=tp(layer.message_layer_2), g_msg2=GatePlan(layer.message_layer_2, device),
                upd1=tp(layer.update_layer_1), g_upd1=GatePlan(layer.update_layer_1, device),
                upd2=tp(layer.update_layer_2)))
        self.hoist_message_layer_1 = True
        # lmax_h = 2 hidden irreps (n x 0e + n x 1o + n x 2e): the edge part of every layer in GEMM form
        self.use_l2_rows = USE_L2_ROWS and all(Lmax2EdgePlan.supported(layer, model.hidden_irreps)
                                               for layer in model.layers)
        if self.use_l2_rows:
            for layer, pl in zip(model.layers, self.layers):
                pl["l2"] = Lmax2EdgePlan(layer, model.hidden_irreps, pl["msg1h"], pl["msg2"], device)
        self.pool1, self.g_pool1 = tp(model.pre_pool1), GatePlan(model.pre_pool1, device)
        self.pool2 = tp(model.pre_pool2)

    @torch.no_grad()
    def forward_edge_list(self, pos, vel, mass, edge_index, B: int, N: int, return_layers: bool = False, x_in=None,
                          node_attr=None):
        """SEGNN.forward (segnn.py:150-189) on an explicit edge list -- the kNN graphs build_graph_with_knn returns for
        num_neighbors < N - 1 (utils/build_fully_connected_graph.py:42-80): gathered message input, one tensor product
        per call, eval BatchNorm per edge, deterministic segment sum over the incoming edges of every node."""
        model = self.model
        if model.training and model.norm == "batch":
            raise NotImplementedError("the generic-irreps path implements eval-mode BatchNorm only")
        D, nodes = model.hidden_irreps.dim, pos.shape[0]
        edge_index = edge_index.to(device=pos.device, dtype=torch.int64).contiguous()
        E = edge_index.shape[1]
        order, ptr = ops.edge_list_csr(edge_index, nodes)
        ea, add = ops.edge_attr_list(pos, mass, edge_index, model.lmax_attr)
        if x_in is None or node_attr is None:
            x_in, attr = ops.prep_list(pos, vel, ea, order, ptr, model.lmax_attr)
        else:  # the graph's own x / node_attr (e.g. O3Transform(use_force_input=True))
            attr = node_attr
        x = self.embed.run(x_in, attr)
        per_layer = [x]
        for layer, pl in zip(model.layers, self.layers):
            inp = torch.empty((E, 2 * D + 2), dtype=torch.float32, device=pos.device)
            with torch.cuda.device(pos.device):
                check(lib.segnn_generic_message_input_list(_p(x), _p(add), _p(edge_index), E, D, 2, _p(inp),
                                                           ops._stream()), "segnn_generic_message_input_list")
            ops._bump()
            m = pl["g_msg1"].run(pl["msg1"].run(inp, ea))
            m = pl["g_msg2"].run(pl["msg2"].run(m, ea))
            if layer.message_norm is not None:  # per edge, as the reference applies it (segnn.py:281-283)
                mul, addc = _bn_eval_columns(layer.message_norm, layer.hidden_irreps)
                m = ops.lincomb(m, None, mul, None, addc)
            agg = ops.segment_reduce(m, order, ptr)
            u = pl["g_upd1"].run(pl["upd1"].run(torch.cat([x, agg], dim=1).contiguous(), attr))
            u = pl["upd2"].run(u, attr)
            x = ops.add3(x, u)
            if layer.feature_norm is not None:
                x = _feature_norm(layer, x, B, N)
            per_layer.append(x)
        h = self.g_pool1.run(self.pool1.run(x, attr))
        pred = self.pool2.run(h, attr)
        return (pred, per_layer) if return_layers else pred

    @torch.no_grad()
    def forward(self, pos, vel, mass, B: int, N: int, return_layers: bool = False, x_in=None, node_attr=None):
        model = self.model
        if model.training and model.norm == "batch":
            raise NotImplementedError("the generic-irreps path implements eval-mode BatchNorm only")
        D = model.hidden_irreps.dim
        if x_in is None or node_attr is None:
            x_in, attr = ops.prep(pos, vel, B, N, model.lmax_attr)
        else:  # the graph's own x / node_attr (e.g. O3Transform(use_force_input=True))
            attr = node_attr
        if not self.use_l2_rows:
            ea, add = ops.edge_attr(pos, mass, B, N, model.lmax_attr)
            E = ea.shape[0]
        x = self.embed.run(x_in, attr)
        per_layer = [x]
        for layer, pl in zip(model.layers, self.layers):
            if self.use_l2_rows:
                agg = pl["l2"].run(x, pos, mass, B, N)
                u = pl["g_upd1"].run(pl["upd1"].run(torch.cat([x, agg], dim=1).contiguous(), attr))
                u = pl["upd2"].run(u, attr)
                x = ops.add3(x, u)
                if layer.feature_norm is not None:
                    x = _feature_norm(layer, x, B, N)
                per_layer.append(x)
                continue
            if self.hoist_message_layer_1:
                m = pl["g_msg1"].run(pl["msg1h"].run(x, ea, add, B, N))
            else:  # the reference's own formulation: gathered message input, tensor product on [E, 2D + 2]
                inp = torch.empty((E, 2 * D + 2), dtype=torch.float32, device=pos.device)
                with torch.cuda.device(pos.device):
                    check(lib.segnn_generic_message_input(_p(x), _p(add), B, N, D, 2, _p(inp), ops._stream()),
                          "segnn_generic_message_input")
                m = pl["g_msg1"].run(pl["msg1"].run(inp, ea))
            m = pl["g_msg2"].run(pl["msg2"].run(m, ea))
            agg = torch.empty((B * N, D), dtype=torch.float32, device=pos.device)
            with torch.cuda.device(pos.device):
                check(lib.segnn_generic_aggregate(_p(m), B, N, D, _p(agg), ops._stream()), "segnn_generic_aggregate")
            ops._bump(2)
            if layer.message_norm is not None:
                # eval-mode BatchNorm is affine per column, so it commutes with the sum over the N - 1 senders:
                # sum_j (mul m_ij + add) = mul sum_j m_ij + (N - 1) add -- applied on [nodes, D], not on [E, D]
                mul, addc = _bn_eval_columns(layer.message_norm, layer.hidden_irreps)
                agg = ops.lincomb(agg, None, mul, None, addc * float(N - 1))
            u = pl["g_upd1"].run(pl["upd1"].run(torch.cat([x, agg], dim=1).contiguous(), attr))
            u = pl["upd2"].run(u, attr)
            x = ops.add3(x, u)
            if layer.feature_norm is not None:
                x = _feature_norm(layer, x, B, N)
            per_layer.append(x)
        h = self.g_pool1.run(self.pool1.run(x, attr))
        pred = self.pool2.run(h, attr)
        return (pred, per_layer) if return_layers else pred
