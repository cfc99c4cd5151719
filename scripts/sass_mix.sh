#!/bin/bash
# SASS instruction mix of the headline kernel variant (k_fused<FASTB=1, DMODE=1 (matrix), EMITC=1 (one-line adptSim dialects)>) and of
# k_deam_stream: bash scripts/sass_mix.sh [lib.so] > profiles/r02_sass_mix.txt
LIB=${1:-gargammel_b200/libgargammel_b200.so}
for K in '_ZN3ggk7k_fusedILb1ELi1ELi1EEEvNS_11FusedParamsE' 'k_deam_stream'; do
  cuobjdump -sass $LIB 2>/dev/null | awk -v k="$K" '/Function :/{f = index($0, k) > 0} f' > /tmp/sass_$$.txt
  N=$(grep -cE '^\s+/\*[0-9a-f]{4,5}\*/' /tmp/sass_$$.txt)
  echo "== $K: $N SASS instructions ($((N * 16)) bytes)"
  grep -E '^\s+/\*[0-9a-f]{4,5}\*/' /tmp/sass_$$.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -28
  echo "-- wide / async memory ops:"
  grep -oE '\b(STG\.E\.128[A-Z.]*|STS\.128|LDS\.128|LDG\.E\.128[A-Z.]*|LDG\.E\.64[A-Z.]*|UBLKCP[A-Z.]*|LDGSTS[A-Z.0-9]*|ATOMG[A-Z.0-9]*|CCTL[A-Z.]*|PRMT|SHF\.R\.W\.U32|IMAD\.WIDE\.U32|REDUX[A-Z.]*)\b' /tmp/sass_$$.txt | sort | uniq -c | sort -rn
  echo
done
rm -f /tmp/sass_$$.txt
