"""Stand-in for ``e3nn.nn`` (e3nn 0.5.1): ``Gate``, ``Activation``, ``BatchNorm`` -- TEST INFRASTRUCTURE, see
../../README.md.  Follows the structure of e3nn's own modules ([batch, sample, mul, repr] reshapes, ``_roll_avg``,
``normalize2mom`` with 1e6 float64 samples from ``torch.Generator().manual_seed(0)``)."""
import torch
import torch.nn as nn

from ..o3 import Irreps

__all__ = ["Gate", "Activation", "BatchNorm", "normalize2mom"]


def _moment(f, n: int) -> torch.Tensor:
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    return f(z).pow(n).mean()


class normalize2mom(nn.Module):
    def __init__(self, f):
        super().__init__()
        with torch.no_grad():
            cst = _moment(f, 2).pow(-0.5).item()
        if abs(cst - 1) < 1e-4:
            self._is_id, self.cst = True, 1.0
        else:
            self._is_id, self.cst = False, cst
        self.f = f

    def forward(self, x):
        return self.f(x) if self._is_id else self.f(x).mul(self.cst)


class Activation(nn.Module):
    def __init__(self, irreps_in, acts):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        assert len(self.irreps_in) == len(acts)
        for (mul, ir), act in zip(self.irreps_in, acts):
            assert act is None or ir.l == 0, "Activation: cannot apply an activation function to a non-scalar input."
        self.acts = nn.ModuleList([normalize2mom(a) if a is not None else None for a in acts])
        self.irreps_out = self.irreps_in

    def forward(self, features):
        out, i = [], 0
        for (mul, ir), act in zip(self.irreps_in, self.acts):
            blk = features.narrow(-1, i, mul * ir.dim)
            out.append(act(blk) if act is not None else blk)
            i += mul * ir.dim
        return torch.cat(out, dim=-1) if len(out) > 1 else out[0]


class Gate(nn.Module):
    def __init__(self, irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated):
        super().__init__()
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = \
            Irreps(irreps_scalars), Irreps(irreps_gates), Irreps(irreps_gated)
        if len(self.irreps_gates) > 0 and self.irreps_gates.lmax > 0:
            raise ValueError(f"Gate scalars must be scalars, instead got irreps_gates = {irreps_gates}")
        if len(self.irreps_scalars) > 0 and self.irreps_scalars.lmax > 0:
            raise ValueError(f"Scalars must be scalars, instead got irreps_scalars = {irreps_scalars}")
        if self.irreps_gates.num_irreps != self.irreps_gated.num_irreps:
            raise ValueError("There are {} irreps in irreps_gated, but a different number ({}) of gate scalars"
                             .format(self.irreps_gated.num_irreps, self.irreps_gates.num_irreps))
        self.act_scalars = Activation(self.irreps_scalars, act_scalars)
        self.act_gates = Activation(self.irreps_gates, act_gates)
        self.irreps_in = self.irreps_scalars + self.irreps_gates + self.irreps_gated
        self.irreps_out = self.irreps_scalars + self.irreps_gated

    def forward(self, features):
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        scalars = features.narrow(-1, 0, ns)
        gates = features.narrow(-1, ns, ng)
        gated = features.narrow(-1, ns + ng, self.irreps_gated.dim)
        scalars = self.act_scalars(scalars)
        if ng:
            gates = self.act_gates(gates)
            # ElementwiseTensorProduct(irreps_gated, irreps_gates): the k-th gate multiplies the k-th gated irrep
            cols, iv, ig = [], 0, 0
            for mul, ir in self.irreps_gated:
                blk = gated.narrow(-1, iv, mul * ir.dim).reshape(*gated.shape[:-1], mul, ir.dim)
                cols.append((blk * gates.narrow(-1, ig, mul).unsqueeze(-1)).reshape(*gated.shape[:-1], mul * ir.dim))
                iv += mul * ir.dim
                ig += mul
            features = torch.cat([scalars] + cols, dim=-1)
        else:
            features = scalars
        return features


class BatchNorm(nn.Module):
    def __init__(self, irreps, eps=1e-5, momentum=0.1, affine=True, reduce="mean", instance=False,
                 normalization="component"):
        super().__init__()
        self.irreps, self.eps, self.momentum, self.affine, self.instance = Irreps(irreps), eps, momentum, affine, instance
        num_scalar = sum(mul for mul, ir in self.irreps if ir.l == 0 and ir.p == 1)
        num_features = self.irreps.num_irreps
        if self.instance:
            self.register_buffer("running_mean", None)
            self.register_buffer("running_var", None)
        else:
            self.register_buffer("running_mean", torch.zeros(num_scalar))
            self.register_buffer("running_var", torch.ones(num_features))
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_scalar))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        assert reduce in ["mean", "max"] and normalization in ["norm", "component"]
        self.reduce, self.normalization = reduce, normalization

    def _roll_avg(self, curr, update):
        return (1 - self.momentum) * curr + self.momentum * update.detach()

    def forward(self, input):
        batch, *size, dim = input.shape
        input = input.reshape(batch, -1, dim)  # [batch, sample, stacked features]
        if self.training and not self.instance:
            new_means, new_vars = [], []
        fields, ix, irm, irv, iw, ib = [], 0, 0, 0, 0, 0
        for mul, ir in self.irreps:
            d = ir.dim
            field = input[:, :, ix: ix + mul * d].reshape(batch, -1, mul, d)  # [batch, sample, mul, repr]
            ix += mul * d
            if ir.l == 0 and ir.p == 1:
                if self.training or self.instance:
                    if self.instance:
                        field_mean = field.mean(1).reshape(batch, mul)
                    else:
                        field_mean = field.mean([0, 1]).reshape(mul)
                        new_means.append(self._roll_avg(self.running_mean[irm: irm + mul], field_mean))
                else:
                    field_mean = self.running_mean[irm: irm + mul]
                irm += mul
                field = field - field_mean.reshape(-1, 1, mul, 1)
            if self.training or self.instance:
                if self.normalization == "norm":
                    field_norm = field.pow(2).sum(3)
                else:
                    field_norm = field.pow(2).mean(3)
                field_norm = field_norm.mean(1) if self.reduce == "mean" else field_norm.max(1).values
                if not self.instance:
                    field_norm = field_norm.mean(0)
                    new_vars.append(self._roll_avg(self.running_var[irv: irv + mul], field_norm))
            else:
                field_norm = self.running_var[irv: irv + mul]
            irv += mul
            field_norm = (field_norm + self.eps).pow(-0.5)
            if self.affine:
                field_norm = field_norm * self.weight[None, iw: iw + mul]
                iw += mul
            field = field * field_norm.reshape(-1, 1, mul, 1)
            if self.affine and (ir.l == 0 and ir.p == 1):
                field = field + self.bias[ib: ib + mul].reshape(mul, 1)
                ib += mul
            fields.append(field.reshape(batch, -1, mul * d))
        assert ix == dim
        if self.training and not self.instance:
            assert irm == self.running_mean.numel() and irv == self.running_var.size(0)
            if len(new_means) > 0:
                torch.cat(new_means, out=self.running_mean)
            if len(new_vars) > 0:
                torch.cat(new_vars, out=self.running_var)
        output = torch.cat(fields, dim=2)
        return output.reshape(batch, *size, dim)
