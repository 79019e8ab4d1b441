#!/bin/bash
# Instruction histogram of the shipped library (CPU box): which Blackwell-native instructions each kernel contains.
# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG (bulk tensor reduction)/UBLKCP, legacy tensor path -> HMMA.
cd "$(dirname "$0")/.."
LIB=instageo-e2e-geospatial-ml_b200/libinstageo_b200.so
echo "# cuobjdump -sass $LIB  (sm_100a), $(date -u +%Y-%m-%d), commit $(git rev-parse --short HEAD)"
cuobjdump -sass $LIB | awk '
  /Function :/ { fn=$3; next }
  /^ +\/\*[0-9a-f]+\*\// {
    op=$2; sub(/;$/,"",op);
    if (op ~ /^@/) { op=$3; sub(/;$/,"",op) }
    base=op; sub(/\..*/,"",base);
    if (base ~ /^(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTMAPF|UBLKCP|HMMA|MUFU|UTCBAR|SYNCS|UTCCP|ACQBULK|ELECT|UCGABAR_ARV|UCGABAR_WAIT)$/) cnt[fn" "op]++;
  }
  END { for (k in cnt) print cnt[k], k }' | sort -k2,2 -k1,1nr | awk '{ if ($2 != last) { print ""; print $2; last=$2 } printf "    %6d  %s\n", $1, $3 }' | c++filt
