#!/usr/bin/env bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md "What proves a
# Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP,
# tcgen05.commit -> UTCBAR, mbarrier -> SYNCS.  Reads the built library; no GPU needed.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
lib="$(ls "${here}"/extending-*_b200/libsegnn_b200.so)"
out="${1:-${here}/profiles/r2_sass_evidence.txt}"
cuobjdump -sass "${lib}" | awk '
  /Function :/ { fn=$3; next }
  { for (m in pat) if ($0 ~ pat[m]) cnt[fn, m]++ }
  BEGIN { pat["UTCHMMA"]="UTCHMMA"; pat["UTCQMMA"]="UTCQMMA"; pat["LDTM"]="LDTM"; pat["STTM"]="STTM"; pat["UBLKCP"]="UBLKCP";
          pat["UTCBAR"]="UTCBAR"; pat["UTCATOM"]="UTCATOMSWS"; pat["SYNCS"]="SYNCS"; pat["HFMA2"]="HFMA2"; pat["FFMA2"]="FFMA2";
          pat["MUFU.TANH"]="MUFU.TANH"; pat["HMMA"]=" HMMA"; pat["RED/ATOM"]="(RED|ATOMG)\\." }
  END {
    n = split("UTCHMMA UTCQMMA LDTM STTM UBLKCP UTCBAR UTCATOM SYNCS HFMA2 FFMA2 MUFU.TANH HMMA RED/ATOM", cols, " ");
    for (key in cnt) { split(key, kk, SUBSEP); fns[kk[1]] = 1 }
    for (fn in fns) {
      line = ""; tot = 0;
      for (i = 1; i <= n; i++) { c = cnt[fn, cols[i]] + 0; if (c) line = line sprintf(" %s=%d", cols[i], c); if (i <= 7) tot += c }
      if (tot) printf "%s:%s\n", fn, line
    }
  }' | sort > "${out}.tmp"
{
  echo "# SASS evidence from $(basename "${lib}") (cuobjdump -sass, $(nvcc --version | tail -2 | head -1))"
  echo "# kernels that contain tcgen05 / TMEM / bulk-copy instructions, with instruction counts"
  while read -r line; do
    fn="${line%%:*}"; echo "$(echo "${fn}" | c++filt | cut -c1-110):${line#*:}"
  done < "${out}.tmp"
  echo "# totals"
  cuobjdump -sass "${lib}" | grep -o -E 'UTCHMMA|LDTM|STTM|UBLKCP|UTCBAR|UTMALDG| HMMA|HGMMA' | sort | uniq -c
} > "${out}"
rm -f "${out}.tmp"
echo "wrote ${out}"
