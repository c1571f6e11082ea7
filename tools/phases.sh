#!/bin/bash
N=${1:-20}; shift
for d in "$@"; do echo -n "DBG=$d EX=${PB200_MAX_EXTRA:-16}: "; PB200_DBG=$d python tools/apply_only.py $N 50 2>&1 | tail -1; done
