#!/bin/bash
# registers / shared memory / spills per kernel from the ptxas logs of the last build:  tools/kinfo.sh [pattern]
cd "$(dirname "$0")/../zstd_b200/csrc"
for f in *.ptxas.log; do
  awk -v pat="${1:-.}" '/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)}
       /Used [0-9]+ registers/ { if (name ~ pat) { spill=""; print substr(name,1,60) " : " $0 } }
       /spill/ { if (name ~ pat && $0 !~ / 0 bytes spill stores, 0 bytes spill loads/) print "   SPILL " $0 }' "$f"
done
