#!/bin/bash
# c3 statistics kernel under different block shapes / buffer counts (CLAIR_TUNE knobs)
for tune in "" "stats_buffers=1" "stats_warps=8,stats_slots=4" "stats_warps=8,stats_slots=4,stats_buffers=1" "stats_warps=15,stats_slots=2" "stats_warps=16,stats_slots=2,stats_blocks_per_sm=1" "stats_warps=10,stats_slots=4" ; do
  CLAIR_TUNE=$tune python scratch/time_pairs.py c3 2>&1 | tail -1
done
