#!/bin/bash
# Evidence run for profiles/: default bench, reference arm, launch lists and ncu --set full captures of the pair kernels.
# Every ncu pass runs only after the same command has exited 0 without ncu.
set -u
out=gpurun_out/prof
mkdir -p $out
python bench.py > $out/bench_default.json 2> $out/bench_default.err || exit 1
python bench.py --impl reference --steps 10 --warmup 3 > $out/bench_reference.json 2> $out/bench_reference.err || exit 1
python scratch/prof_train.py train 3 > $out/train_plain.log 2>&1 || exit 1
python scratch/prof_train.py lin 2 > $out/lin_plain.log 2>&1 || exit 1
python scratch/prof_c5.py 2 > $out/c5_plain.log 2>&1 || exit 1
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --csv --log-file $out/train_step_launches.csv python scratch/prof_train.py train 3 > $out/ncu_train_list.log 2>&1
ncu --metrics $M --clock-control none -k regex:pair --csv --log-file $out/c5_step_launches.csv python scratch/prof_c5.py 1 > $out/ncu_c5_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'pair_stats|pair_grad|pair_upstream' --launch-skip 3 -c 3 -o $out/train_kernels python scratch/prof_train.py train 3 > $out/ncu_train_full.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:pair_stats -c 1 -o $out/pair_stats_c3 python scratch/prof_train.py lin 1 > $out/ncu_lin_full.log 2>&1
ls -la $out
