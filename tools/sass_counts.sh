#!/bin/bash
# Per-kernel counts of the Blackwell-native SASS instructions in the shipped library (B200_PROFILING.md: tcgen05.mma ->
# UTCHMMA, tcgen05.ld -> LDTM, TMA bulk copy -> UBLKCP, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS).  No GPU needed.
#   bash tools/sass_counts.sh > profiles/sass_r02.txt
LIB=${1:-yet-another-nerf_b200/libyanerf_b200.so}
echo "# cuobjdump -sass $LIB ($(date -u +%Y-%m-%d), nvcc $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# kernel UTCHMMA LDTM UBLKCP UTCBAR SYNCS HMMA(legacy)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn = $3 }
  { for (i = 1; i <= NF; ++i) { m = $i; sub(/\..*/, "", m);
      if (m == "UTCHMMA" || m == "LDTM" || m == "UBLKCP" || m == "UTCBAR" || m == "SYNCS" || m == "HMMA") c[fn, m]++ } ; seen[fn] = 1 }
  END { for (fn in seen) if (fn != "") printf "%s %d %d %d %d %d %d\n", fn, c[fn, "UTCHMMA"], c[fn, "LDTM"], c[fn, "UBLKCP"], c[fn, "UTCBAR"], c[fn, "SYNCS"], c[fn, "HMMA"] }' |
  sort | while read fn a b c d e f; do echo "$(echo $fn | c++filt | cut -c1-90) | $a $b $c $d $e $f"; done
