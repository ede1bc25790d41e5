#!/bin/bash
# Counts the SASS mnemonics that show which hardware paths the built kernels use (no GPU needed):
#   UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor
#   load / store, UTMAPF = TMA prefetch into L2, SYNCS = mbarrier operations, HMMA = legacy mma.sync (must be 0).
# usage: bash scripts/sass_evidence.sh > profiles/r1_sass_evidence.md   (after building liboasr.so)
cd "$(dirname "$0")/../omnilingual-asr_b200/csrc" || exit 1
echo "# SASS evidence (cuobjdump -sass of the sm_100a objects behind liboasr.so)"
echo
echo "| object | kernels | UTCHMMA | UTCBAR | LDTM | STTM | UTMALDG | UTMASTG | UTMAPF | SYNCS | HMMA | FFMA2 | MUFU.EX2 |"
echo "|---|---|---|---|---|---|---|---|---|---|---|---|---|"
for f in gemm_tcgen05 attention_v4 attention_v7 norm_conv0 decode tp_fused resample; do
  s=$(cuobjdump -sass build/$f.o 2>/dev/null)
  c() { grep -c "$1" <<<"$s"; }
  echo "| $f | $(c 'Function :') | $(c UTCHMMA) | $(c UTCBAR) | $(c LDTM) | $(c STTM) | $(c UTMALDG) | $(c UTMASTG) | $(c UTMAPF) | $(c SYNCS) | $(c ' HMMA') | $(c FFMA2) | $(c 'MUFU.EX2') |"
done
