"""CPU tests pinning the oracle: closed forms, constants, hand-written graph enumerations, symmetry properties and
the committed golden fixture (SURVEY 8(c) list of golden vectors the build must create itself)."""
import math
import os

import pytest
import torch

from oracle import segnn_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "segnn_small.pt")


def test_edge_enumeration_by_hand():
    # nonzero(~eye(N)) row-major: source ascending, then target ascending, graph offset g*N
    assert O.fully_connected_edge_index(1, 2).tolist() == [[0, 1], [1, 0]]
    assert O.fully_connected_edge_index(2, 3).tolist() == [
        [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5], [1, 2, 0, 2, 0, 1, 4, 5, 3, 5, 3, 4]]
    ei = O.fully_connected_edge_index(3, 5)
    assert ei.shape == (2, 60)
    assert ei[:, :4].tolist() == [[0, 0, 0, 0], [1, 2, 3, 4]]
    assert ei[:, 20:24].tolist() == [[5, 5, 5, 5], [6, 7, 8, 9]]
    assert ei[:, -1].tolist() == [14, 13]
    with pytest.raises(ValueError):
        O.build_graph_with_knn(None, 1, 5, None, 5)


def test_normalize2mom_constants():
    assert abs(O.normalize2mom_const("silu") - 1.6791767923989418) < 1e-12
    assert abs(O.normalize2mom_const("sigmoid") - 1.8467055342154763) < 1e-12


def test_wigner_identities():
    eye = torch.eye(3, dtype=torch.float64)
    assert torch.allclose(O.wigner_3j(0, 0, 0), torch.ones(1, 1, 1, dtype=torch.float64))
    assert torch.allclose(O.wigner_3j(0, 1, 1)[0], eye / math.sqrt(3))
    assert torch.allclose(O.wigner_3j(1, 0, 1)[:, 0, :], eye / math.sqrt(3))
    assert torch.allclose(O.wigner_3j(1, 1, 0)[:, :, 0], eye / math.sqrt(3))
    c = O.wigner_3j(1, 1, 2)
    s10, s30 = 1 / math.sqrt(10), 1 / math.sqrt(30)
    expect = torch.zeros(3, 3, 5, dtype=torch.float64)
    expect[0, 2, 0] = expect[2, 0, 0] = s10
    expect[0, 1, 1] = expect[1, 0, 1] = s10
    expect[0, 0, 2], expect[1, 1, 2], expect[2, 2, 2] = -s30, 2 * s30, -s30
    expect[1, 2, 3] = expect[2, 1, 3] = s10
    expect[0, 0, 4], expect[2, 2, 4] = -s10, s10
    assert torch.allclose(c, expect, atol=1e-12)
    assert torch.allclose(O.wigner_3j(2, 1, 1), c.permute(2, 1, 0), atol=1e-12)


# ---- third-party anchors (SymPy): the oracle cannot be pinned against e3nn itself (not installable here), but the two
# conventions everything else hangs on -- the real spherical-harmonic basis and the coupling tensors in that basis -- can
# be pinned against an independent implementation of the published definitions. -------------------------------------
COUPLINGS = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 2), (2, 0, 2), (2, 1, 1), (2, 2, 0), (2, 2, 2)]


def test_wigner_3j_equals_sympy_real_gaunt_coefficients():
    """Every coupling the SEGNN tensor products use (l <= 2 features, l <= 1 attributes; l1 + l2 + l3 even) equals the
    normalised real Gaunt coefficient  int Z_{l1 m1} Z_{l2 m2} Z_{l3 m3} dOmega  of SymPy -- same basis order m = -l..l,
    same sign, unit Frobenius norm -- for the oracle's table and for the package's own (cg.py)."""
    sympy_wigner = pytest.importorskip("sympy.physics.wigner")
    import numpy as np
    import segnn_b200.cg as cg
    for l1, l2, l3 in COUPLINGS:
        g = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
        for m1 in range(-l1, l1 + 1):
            for m2 in range(-l2, l2 + 1):
                for m3 in range(-l3, l3 + 1):
                    g[l1 + m1, l2 + m2, l3 + m3] = float(sympy_wigner.real_gaunt(l1, l2, l3, m1, m2, m3))
        g /= np.linalg.norm(g)
        assert np.abs(O.wigner_3j(l1, l2, l3).numpy() - g).max() < 1e-12, (l1, l2, l3)
        assert np.abs(cg.real_wigner_3j(l1, l2, l3) - g).max() < 1e-12, (l1, l2, l3)


def _sympy_real_harmonics(point):
    """Real spherical harmonics l <= 2 of SymPy (Znm) at a unit vector given in the oracle's axis convention: e3nn's
    l = 1 basis is (x, y, z) where the textbook one (m = -1, 0, 1) is (y, z, x), i.e. the frames differ by a cyclic
    relabelling of the axes; SymPy's Znm carries the Condon-Shortley sign (-1)^m for m > 0 and -1 for m < 0."""
    import sympy
    xs, ys, zs = point[2], point[0], point[1]
    theta, phi = math.acos(max(-1.0, min(1.0, zs))), math.atan2(ys, xs)
    out = []
    for l in range(3):
        for m in range(-l, l + 1):
            sign = 1.0 if m == 0 else ((-1.0) ** m if m > 0 else -1.0)
            out.append(sign * float(sympy.re(sympy.N(sympy.Znm(l, m, theta, phi).expand(func=True)))))
    return out


def test_spherical_harmonics_equal_sympy_real_harmonics():
    pytest.importorskip("sympy")
    gen = torch.Generator().manual_seed(3)
    pts = torch.randn(5, 3, generator=gen, dtype=torch.float64)
    pts = pts / pts.norm(dim=1, keepdim=True)
    sh = O.spherical_harmonics(2, pts, normalize=True)
    for k in range(pts.shape[0]):
        ref = torch.tensor(_sympy_real_harmonics([float(v) for v in pts[k]]), dtype=torch.float64)
        assert float((sh[k] - ref).abs().max()) < 1e-12


def test_wigner_3j_is_the_gaunt_integral_of_the_oracle_harmonics():
    """Basis consistency without any third party: integrating the product of three oracle harmonics over the sphere
    (Gauss-Legendre in cos(theta) x uniform in phi: exact for these polynomials) reproduces the oracle's coupling tensor
    up to its normalisation, so the harmonics (edge / node attributes) and the couplings live in the same basis."""
    import numpy as np
    nodes, weights = np.polynomial.legendre.leggauss(8)
    phis = (np.arange(16) + 0.5) * (2 * math.pi / 16)
    ct, ph = np.meshgrid(nodes, phis, indexing="ij")
    st = np.sqrt(1 - ct ** 2)
    # the oracle's axes: z_e = std y, ... any right-handed frame works for the integral; use x = st cos, y = st sin, z = ct
    vec = torch.tensor(np.stack([st * np.cos(ph), st * np.sin(ph), ct], axis=-1).reshape(-1, 3))
    wq = torch.tensor((weights[:, None] * np.ones_like(ph) * (2 * math.pi / 16)).reshape(-1))
    sh = O.spherical_harmonics(2, vec, normalize=True)
    off = {0: 0, 1: 1, 2: 4}
    for l1, l2, l3 in COUPLINGS:
        a, b, c = (sh[:, off[l]: off[l] + 2 * l + 1] for l in (l1, l2, l3))
        g = torch.einsum("p,pi,pj,pk->ijk", wq, a, b, c)
        g = g / g.norm()
        assert float((g - O.wigner_3j(l1, l2, l3)).abs().max()) < 1e-12, (l1, l2, l3)


def test_tensor_product_closed_forms():
    # net coefficient of every path is sqrt(2 lo + 1) * C: identities and dot/sqrt(3)
    torch.manual_seed(0)
    tp = O.O3TensorProduct("3x0e+3x1o", "2x0e+2x1o", "1x0e+1x1o")
    x = torch.randn(5, 12, dtype=torch.float64)
    a = torch.randn(5, 4, dtype=torch.float64)
    out = tp(x, a)
    wss, wsv, wvv, wvs = [w[:, 0, :] for w in tp.tp.weight_views()]
    s, v = x[:, :3], x[:, 3:].reshape(5, 3, 3)
    a0, a1 = a[:, :1], a[:, 1:]
    exp0 = a0 * (s @ wss) + ((v * a1[:, None, :]).sum(-1) / math.sqrt(3)) @ wvs + tp.biases
    exp1 = (s @ wsv)[:, :, None] * a1[:, None, :] + a0[:, :, None] * torch.einsum("zuk,uw->zwk", v, wvv)
    assert torch.allclose(out[:, :2], exp0, atol=1e-12)
    assert torch.allclose(out[:, 2:].reshape(5, 2, 3), exp1, atol=1e-12)


def test_parameter_counts_and_irreps():
    for H, l, L, count, irr in [(64, 1, 4, 148256, "32x0e+32x1o"), (192, 1, 6, 1947552, "96x0e+96x1o"),
                                (128, 1, 6, 868288, "64x0e+64x1o"), (192, 2, 6, 2053709, "73x0e+73x1o+73x2e")]:
        m = O.SEGNN(hidden_features=H, lmax_h=l, num_layers=L)
        assert str(m.hidden_irreps) == irr
        assert sum(p.numel() for p in m.parameters()) == count


def _random_rotation(seed):
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def _rotate_features(x, n, R):
    rows = x.shape[0]
    return torch.cat([x[:, :n], (x[:, n:].reshape(rows, n, 3) @ R.T).reshape(rows, 3 * n)], dim=1)


@pytest.mark.parametrize("improper", [False, True])
def test_layer_equivariance(improper):
    """Rotation (and inversion) equivariance of one SEGNNLayer on features downstream of the pos.mean(1) quirk."""
    torch.manual_seed(0)
    n, B, N = 6, 2, 5
    h = O.Irreps(f"{n}x0e+{n}x1o")
    layer = O.SEGNNLayer(h, h, h, "1x0e+1x1o", "1x0e+1x1o", norm="batch", additional_message_irreps="2x0e").eval()
    O.perturb_bn_buffers(layer)
    R = _random_rotation(3) * (-1.0 if improper else 1.0)
    pos, vel, mass = O.synthetic_system(B, N, seed=2)
    pos, vel, mass = pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1)
    x = torch.randn(B * N, 4 * n, dtype=torch.float64)
    outs = []
    for p, v, xx in [(pos, vel, x), (pos @ R.T, vel @ R.T, _rotate_features(x, n, R))]:
        g = O.make_graph(p, v, mass, B, N)
        g.node_attr[:, 0] = 1.0
        with torch.no_grad():
            outs.append(layer(xx, g.edge_index, g.edge_attr, g.node_attr, None, g.additional_message_features))
    assert torch.allclose(_rotate_features(outs[0], n, R), outs[1], atol=1e-10)


def test_permutation_equivariance():
    torch.manual_seed(0)
    B, N = 2, 6
    model = O.SEGNN(hidden_features=16, num_layers=2).eval()
    O.perturb_bn_buffers(model)
    pos, vel, mass = O.synthetic_system(B, N, seed=4)
    perm = torch.stack([torch.randperm(N) for _ in range(B)])
    gather = lambda t: torch.stack([t[b][perm[b]] for b in range(B)])
    with torch.no_grad():
        a = model(O.make_graph(pos.reshape(-1, 3), vel.reshape(-1, 3), mass.reshape(-1, 1), B, N)).reshape(B, N, 6)
        b = model(O.make_graph(gather(pos).reshape(-1, 3), gather(vel).reshape(-1, 3), gather(mass).reshape(-1, 1),
                               B, N)).reshape(B, N, 6)
    assert torch.allclose(gather(a), b, atol=1e-10)


def test_batchnorm_train_eval_agree_after_momentum_steps():
    torch.manual_seed(0)
    bn = O.BatchNorm("4x0e+4x1o")
    x = torch.randn(4096, 16, dtype=torch.float64) * 1.7 + 0.3
    bn.train()
    for _ in range(200):
        y_train = bn(x)
    bn.eval()
    assert torch.allclose(bn(x), y_train, atol=1e-6)
    assert torch.allclose(y_train[:, :4].mean(0), torch.zeros(4, dtype=torch.float64), atol=1e-9)


def test_bn_fold_through_sum_identity():
    """sum_j BN(m_j) == mul * sum_j m_j + deg * (bias - mean*mul): the rewrite the fused edge kernel uses."""
    torch.manual_seed(0)
    bn = O.BatchNorm("3x0e+3x1o").eval()
    O.perturb_bn_buffers(bn)
    m = torch.randn(7, 12, dtype=torch.float64)
    direct = bn(m).sum(0)
    mul = bn.weight * (bn.running_var + bn.eps).rsqrt()
    folded_s = mul[:3] * m[:, :3].sum(0) + 7 * (bn.bias - bn.running_mean * mul[:3])
    folded_v = (m[:, 3:].sum(0).reshape(3, 3) * mul[3:, None]).reshape(9)
    assert torch.allclose(direct, torch.cat([folded_s, folded_v]), atol=1e-12)


def test_oracle_matches_golden_fixture():
    gold = torch.load(GOLDEN)
    cfg = gold["config"]
    model = O.SEGNN(hidden_features=cfg["hidden_features"], num_layers=cfg["num_layers"]).eval()
    model.load_state_dict(gold["state_dict"])
    B, N = cfg["B"], cfg["N"]
    g = O.make_graph(gold["pos"].reshape(-1, 3), gold["vel"].reshape(-1, 3), gold["mass"].reshape(-1, 1), B, N)
    assert torch.equal(g.edge_index, gold["edge_index"])
    assert torch.allclose(g.edge_attr, gold["edge_attr"], atol=1e-14)
    assert torch.allclose(g.node_attr, gold["node_attr"], atol=1e-14)
    with torch.no_grad():
        pred, layers = model(g, return_layers=True)
        loc, vel = O.rollout(model, gold["pos"], gold["vel"], gold["mass"], steps=3)
    assert torch.allclose(pred, gold["pred"], atol=1e-12)
    for a, b in zip(layers, gold["layers"]):
        assert torch.allclose(a, b, atol=1e-12)
    assert torch.allclose(loc, gold["rollout_loc"], atol=1e-11)
    assert torch.allclose(vel, gold["rollout_vel"], atol=1e-11)
