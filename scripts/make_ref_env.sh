#!/usr/bin/env bash
# Tries to install the reference's un-vendored third-party stack into baseline/_ref (git-ignored, travels with gpurun)
# so that oracle/ref_loader.py reports kind = "reference" instead of "reference+shims".
#   e3nn==0.5.1  torch_geometric==2.6.1  torch_scatter==2.1.2   (/root/reference/requirements.txt:9,21,22)
# No index is reachable from the build container or the GPU boxes, so this only succeeds where wheels are available
# (/opt/wheelhouse or a reachable index).  The log is kept either way.
set -uo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
log="${1:-${here}/profiles/r2_ref_env_attempt.log}"
mkdir -p "${here}/baseline/_ref"
{
  echo "== $(date -u +%FT%TZ) host $(hostname) =="
  python -c 'import torch; print("torch", torch.__version__)'
  echo "-- wheelhouse candidates:"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'e3nn|geometric|scatter|opt_einsum' || echo "(none)"
  echo "-- offline install (wheelhouse only)"
  python -m pip install --no-index --find-links /opt/wheelhouse --target "${here}/baseline/_ref" \
      e3nn==0.5.1 torch_geometric==2.6.1 torch_scatter==2.1.2 2>&1 | tail -5
  echo "-- index install (10 s timeout, 0 retries)"
  python -m pip install --timeout 10 --retries 0 --target "${here}/baseline/_ref" --no-deps \
      e3nn==0.5.1 torch_geometric==2.6.1 opt_einsum_fx opt_einsum 2>&1 | tail -5
  echo "-- the reference itself is not a package (pyproject.toml only configures pyright):"
  python -m pip install --no-index --no-build-isolation --no-deps --target "${here}/baseline/_ref" /root/reference 2>&1 | tail -3
  echo "-- result"
  PYTHONPATH="${here}/baseline/_ref" python -c 'import e3nn, torch_geometric, torch_scatter; print("real packages import: kind=reference")' 2>&1 | tail -1
} | tee "${log}"
