#!/bin/bash
# SASS evidence that libdm_b200.so is Blackwell-native (B200_PROFILING.md mnemonics): counts of the tcgen05 / TMEM / TMA
# instructions per kernel.   tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=${1:-disentangle_mlp_b200/lib/libdm_b200.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%F))"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3 }
  /UTCHMMA|UTCQMMA|UTCMMA/ { mma[fn]++ }
  /UTCHMMA.2CTA|UTCMMA.2CTA/ { mma2[fn]++ }
  /LDTM/ { ldtm[fn]++ }
  /UTMALDG/ { tmald[fn]++ }
  /UTMASTG/ { tmast[fn]++ }
  /UTMAREDG/ { tmared[fn]++ }
  /UTCBAR/ { utcbar[fn]++ }
  /SYNCS/ { syncs[fn]++ }
  /RED\.E|REDG|RED\./ { red[fn]++ }
  /SHFL/ { shfl[fn]++ }
  END {
    printf "%-70s %8s %8s %6s %8s %8s %9s %7s %6s %5s %5s\n", "kernel", "UTC*MMA", "(.2CTA)", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "RED", "SHFL"
    for (f in syncs) if (mma[f] || tmald[f]) printf "%-70s %8d %8d %6d %8d %8d %9d %7d %6d %5d %5d\n", substr(f,1,70), mma[f], mma2[f], ldtm[f], tmald[f], tmast[f], tmared[f], utcbar[f], syncs[f], red[f], shfl[f]
  }'
echo
echo "# totals over the library"
for m in UTCHMMA UTCHMMA.2CTA LDTM UTMALDG UTMASTG UTMAREDG UTCBAR; do
  printf "%-14s %d\n" $m $(cuobjdump -sass "$LIB" | grep -c "$m")
done
echo "# tf32 MMA kind:"; cuobjdump -sass "$LIB" | grep -o "UTC[A-Z]*MMA[.A-Z0-9_]*" | sort | uniq -c
