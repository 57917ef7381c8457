#!/bin/bash
# SASS evidence of the Blackwell-native path: mnemonic counts of the shipped library (profiles/rNN_sass_summary.txt)
LIB=${1:-flashmoe_b200/libflashmoe_b200.so}
OUT=${2:-profiles/r02_sass_summary.txt}
SASS=$(mktemp)
/usr/local/cuda/bin/cuobjdump -sass "$LIB" > "$SASS"
{
  echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), $(sha256sum "$LIB" | cut -c1-16))"
  grep -E "^\s*(arch|Function)" "$SASS" | sed 's/^\s*//' | sort | uniq -c
  echo "# mnemonic counts (tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, tcgen05.commit -> UTCBAR,"
  echo "#                  bf16x8 reduction -> REDG.E.ADD.BF16x8 ..., legacy mma.sync would be HMMA)"
  for m in UTCHMMA UTCHMMA.2CTA LDTM UTMALDG UTMALDG.2D.2CTA UBLKCP UTCBAR UTCBAR.2CTA.MULTICAST UTMAPF REDG.E.ADD SYNCS.ARRIVE SYNCS.PHASECHK MATCH.ANY; do
    printf "%-24s %s\n" "$m" "$(grep -c -F "$m" "$SASS")"
  done
  printf "%-24s %s\n" "HMMA (legacy mma.sync)" "$(grep -cE "[^A-Z]HMMA" "$SASS")"
  printf "%-24s %s\n" "IMMA / QMMA (legacy)" "$(grep -cE "[^A-Z](IMMA|QMMA|HGMMA)" "$SASS")"
  echo "# the distinct tensor / TMA / reduction instruction forms"
  grep -oE "(UTCHMMA|LDTM|UTMALDG|UBLKCP|UTCBAR|UTMAPF|REDG\.E\.ADD|UTCATOMSWS)[A-Za-z0-9_.]*" "$SASS" | sort | uniq -c | sort -rn
} > "$OUT"
rm -f "$SASS"
cat "$OUT"
