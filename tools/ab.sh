#!/bin/bash
# A/B builds of libwost.so on one box: tools/ab.sh "<lib> <lib> ..." scenario ...   (two rounds, interleaved)
LIBS=$1; shift
for rep in 1 2; do for L in $LIBS; do for s in "$@"; do echo -n "$(basename $L) "; WOST_LIB=$L python tools/run_one.py $s 4 | tail -1; done; done; done
