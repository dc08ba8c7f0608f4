"""Times segnn_gemm_tn_tf32x3 / segnn_gemm_tf32x3 alone on the configuration-4 shapes (for ncu and CUDA-event timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segnn_b200 as S

rows = 1_000_000
cases = [("tn", rows, 128, 192), ("tn", 3 * rows, 64, 64), ("tng", rows, 192, 192), ("nn", rows, 128, 192), ("nn", 3 * rows, 64, 64),
         ("nn", rows, 192, 128)]
for kind, r, a, b in cases:
    if kind in ("tn", "tng"):
        x, y = torch.randn(r, a, device="cuda"), torch.randn(r, b, device="cuda")
        fn = (lambda: S.ops.gemm_tn_tf32x3(x, y, groups=3)) if kind == "tng" else (lambda: S.ops.gemm_tn_tf32x3(x, y))
        nbytes = 4 * r * (a + b)
    else:
        x, w = torch.randn(r, a, device="cuda"), torch.randn(a, b, device="cuda")
        out = torch.empty(r, b, device="cuda")
        fn = lambda: S.ops.gemm_tf32x3(x, w, out=out)
        nbytes = 4 * r * (a + b)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{kind} rows={r} {a}x{b}: {ms:.3f} ms, {nbytes / ms / 1e9:.2f} TB/s algorithmic")
    del x
