#!/bin/bash
N=${1:-20}
for d in 0 1 2 3 4 8 7 15; do echo -n "DBG=$d: "; PB200_DBG=$d python tools/apply_only.py $N 50 2>&1 | tail -1; done
