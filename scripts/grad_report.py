"""Prints per-parameter gradient errors of the CUDA training path vs float64 autograd through the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as T

for case in [(64, 2, 4, 5, True), (64, 2, 4, 5, False), (192, 1, 2, 20, True), (128, 2, 1, 33, True), (50, 1, 3, 6, True)]:
    om, m, ref, pred, lr, l = T._grad_case(*case)
    print(case, "pred rel", T.rel(pred, ref), "loss", lr, l)
    for (k, a), (_, b) in zip(om.named_parameters(), m.named_parameters()):
        scale = float(a.grad.abs().max()); err = float((a.grad - b.grad.double().cpu()).abs().max())
        print(f"   {k:45s} scale {scale:9.3e} err {err:9.3e} rel {err / max(scale, 1e-30):8.2e}")
