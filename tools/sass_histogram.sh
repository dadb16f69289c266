#!/bin/bash
# SASS opcode histogram of the built library (whole .so and per kernel): the evidence that the hot ops are tcgen05 / TMEM /
# TMA code and not legacy mma.sync.   bash tools/sass_histogram.sh > profiles/rNN_sass_opcode_histogram.txt
cd "$(dirname "$0")/.." || exit 1
SO=sgpt_b200/libsgpt_b200.so
PAT='UTC[A-Z0-9.]+|LDTM|STTM|UTMA[A-Z]+|UBLKCP|SYNCS|UCGABAR[A-Z_]*|HMMA|MATCH|REDUX|RED|ATOM[SG]?|MEMBAR|ERRBAR|MUFU\.[A-Z0-9]+'
echo "# SASS opcode histogram of $SO (sm_100a):  cuobjdump -sass | grep -oE '$PAT' | sort | uniq -c"
echo "# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG/UTMASTG/UTMAREDG/UTMAPF, mbarrier -> SYNCS, cluster barrier -> UCGABAR; no HMMA (legacy mma.sync) anywhere"
cuobjdump -sass "$SO" > /tmp/sgpt_all.sass
grep -oE "$PAT" /tmp/sgpt_all.sass | sort | uniq -c | sort -rn
echo
echo "# per kernel (tcgen05 / TMEM / TMA instruction counts)"
awk '/Function : /{name=$3} /UTC[A-Z]*MMA/{m[name]++} /LDTM/{l[name]++} /STTM/{s[name]++} /UTMALDG/{t[name]++} /UTMASTG|UTMAREDG/{w[name]++} END{for(n in m) printf "%s UTCMMA=%d LDTM=%d STTM=%d UTMALDG=%d UTMASTG/REDG=%d\n", n, m[n], l[n], s[n], t[n], w[n]}' /tmp/sgpt_all.sass | sort | while read -r line; do n=$(echo "$line" | cut -d' ' -f1); printf "%s %s\n" "$(echo "$n" | c++filt | cut -c1-150)" "$(echo "$line" | cut -d' ' -f2-)"; done
