#!/bin/bash
# cuobjdump -sass opcode counts per object of the in-tree build (tcgen05 / TMEM / TMA / MUFU evidence): tools/sass_summary.sh > profiles/rNN_sass_summary.txt
cd "$(dirname "$0")/../rho-diffusion_b200/csrc/build" || exit 1
for o in *.o; do
  echo "== ${o%.o}"
  cuobjdump -sass "$o" 2>/dev/null | grep -oE '\b(UTCHMMA|UTCQMMA|UTCBAR(\.MULTICAST)?|UTMALDG\.[0-9]D(\.MULTICAST)?|UTMASTG\.[0-9]D|UTCATOMSWS\.[A-Z_.]+|LDTM|STTM|UBLKCP\.[A-Z.]+|SYNCS|ELECT|MUFU\.(EX2|TANH|RCP|RSQ)|FFMA2|FADD2|FMUL2|FMNMX3)\b' \
    | sort | uniq -c | sort -rn | awk '{printf "%7d %s;", $1, $2} END {print ""}'
done
