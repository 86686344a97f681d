#!/bin/bash
# A/B two builds of libwost.so on one box: tools/ab.sh <lib_a> <lib_b> [scenario ...]
A=$1; B=$2; shift 2
for rep in 1 2; do for L in $A $B; do for s in "$@"; do echo -n "$(basename $L) "; WOST_LIB=$L python tools/run_one.py $s 4 | tail -1; done; done; done
