#!/bin/bash
# Counts the Blackwell-specific SASS mnemonics per kernel in slimt_b200/libslimt_b200.so (run where the library was
# built; no GPU needed):  UTCIMMA = tcgen05.mma kind::i8, UTCHMMA = kind::f16/tf32, UTMALDG = TMA tile load,
# LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, and the legacy IMMA/HMMA (mma.sync) that must stay at zero.
#   tools/sass_summary.sh > profiles/sass_summary.txt
set -e
cd "$(dirname "$0")/.."
LIB=slimt_b200/libslimt_b200.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  $LIB  ($(stat -c %s $LIB) bytes)  cuobjdump -sass, sm_100a"
echo "# ldd: $(ldd $LIB | awk '{print $1}' | tr '\n' ' ')"
cuobjdump -sass $LIB | awk '
  /Function :/ { fn=$3 }
  { for (i = 1; i <= NF; i++) { m=$i; sub(/\..*/, "", m);
      if (m=="UTCIMMA"||m=="UTCHMMA"||m=="UTCQMMA"||m=="UTMALDG"||m=="LDTM"||m=="STTM"||m=="UTCBAR"||m=="IMMA"||m=="HMMA"||m=="UTMAPF"||m=="ELECT") c[fn" "m]++ } }
  END { for (k in c) print k, c[k] }' | sort | c++filt | awk '
  { fn=$1; for (i=2;i<NF-1;i++) fn=fn" "$i; m=$(NF-1); n=$NF; if (fn!=last) { if (last!="") print ""; printf "%s\n   ", fn; last=fn } printf " %s=%s", m, n }
  END { print "" }'
echo
echo "# totals"
cuobjdump -sass $LIB | grep -oE "\b(UTCIMMA|UTCHMMA|UTMALDG|LDTM|STTM|UTCBAR|IMMA|HMMA)\b" | sort | uniq -c
