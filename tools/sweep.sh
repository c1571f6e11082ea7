#!/bin/bash
N=${1:-20}
for cfg in "1 1 11 16 3" "0 1 11 16 3" "1 1 11 0 3" "0 1 11 0 3" "1 1 11 0 2" "1 1 12 0 3" "1 1 12 16 3" "1 1 11 16 2"; do
  set -- $cfg
  echo -n "PIPE=$1 PDL=$2 TB=$3 EX=$4 RB=$5: "
  PB200_PIPE=$1 PB200_PDL=$2 PB200_TILE_BITS=$3 PB200_MAX_EXTRA=$4 PB200_REG_BITS=$5 python tools/apply_only.py $N 50 2>&1 | tail -1
done
