#!/bin/bash
# Per-kernel tally of the Blackwell-specific SASS in libttr_b200.so (tcgen05 MMA / TMEM / TMA / cluster instructions)
# -> profiles/<name>.md.  Usage: tools/sass_tally.sh [profiles/r2_sass_tally.md]
OUT=${1:-profiles/r2_sass_tally.md}
SO=twotowermlretrieval_b200/libttr_b200.so
{
echo "# SASS tally of \`libttr_b200.so\` (sm_100a), per kernel"
echo
echo "\`cuobjdump -sass $SO\`, instructions counted per kernel ($(date -u +%Y-%m-%d), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))."
echo "UTCHMMA = tcgen05.mma (\`.2CTA\` = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tile load/store/reduce,"
echo "UBLKCP = bulk copy, UTCBAR = tcgen05.commit, UCGABAR = cluster barrier, SYNCS = mbarrier ops."
echo
echo "| kernel | UTCHMMA | of which .2CTA | LDTM | STTM | UTMALDG | UTMASTG | UTMAREDG | UBLKCP | UTCBAR | UCGABAR | SYNCS | total instrs |"
echo "|---|---|---|---|---|---|---|---|---|---|---|---|---|"
cuobjdump -sass $SO 2>/dev/null | awk '
  /Function : / { if (name != "") flush(); name=$3; for (k in c) delete c[k]; tot=0; next }
  /^[ \t]+\/\*[0-9a-f]+\*\// { tot++; line=$0;
     if (line ~ /UTCHMMA/) { c["mma"]++; if (line ~ /2CTA/) c["mma2"]++ }
     if (line ~ /LDTM/) c["ldtm"]++; if (line ~ /STTM/) c["sttm"]++;
     if (line ~ /UTMALDG/) c["ldg"]++; if (line ~ /UTMASTG/) c["stg"]++; if (line ~ /UTMAREDG/) c["red"]++;
     if (line ~ /UBLKCP/) c["blk"]++; if (line ~ /UTCBAR/) c["bar"]++; if (line ~ /UCGABAR/) c["cga"]++; if (line ~ /SYNCS/) c["syncs"]++ }
  function flush() { if (c["mma"]+c["ldtm"]+c["ldg"]+c["stg"]+c["red"]+c["blk"]+c["cga"] > 0)
     printf "| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n", name, c["mma"], c["mma2"], c["ldtm"], c["sttm"], c["ldg"], c["stg"], c["red"], c["blk"], c["bar"], c["cga"], c["syncs"], tot }
  END { flush() }' | while IFS= read -r l; do
    m=$(echo "$l" | sed -n 's/^| `\([^`]*\)`.*/\1/p'); d=$(echo "$m" | c++filt | sed 's/(.*//' | cut -c1-90); echo "$l" | sed "s|\`$m\`|\`$d\`|"; done
echo
echo "Library dependencies (\`ldd\`): $(ldd $SO | grep -o 'lib[a-z_]*\.so[.0-9]*' | sort -u | tr '\n' ' ')"
} > $OUT
echo wrote $OUT; head -30 $OUT | cut -c1-200
