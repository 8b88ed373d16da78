#!/bin/bash
run() { echo "== lib=$1 tune=$2"; CLAIR_B200_LIB=$PWD/scratch/bin/libclair_$1.so CLAIR_TUNE=$2 python scratch/time_pairs.py c3 2>&1 | tail -1; }
run nohoist ""
run nohoist16 ""
run nohoist128 ""
run nohoist "stats_buffers=1"
