#!/bin/bash
# usage (GPU box): tools/profile_round.sh TAG   ->  gpurun_out/{launches_TAG.csv, prof_TAG_<workload>.ncu-rep, ...}
# The ncu evidence of a round: the launch list of the default bench command, and one `--set full` capture of the encode
# kernel per BASELINE workload (+ the histogram kernel).  Each command has run to completion without ncu first.
tag=${1:-r2}
set -x
python bench.py --steps 3 --warmup 3 --no-cpu --no-per-config > gpurun_out/bench_${tag}_plain.json 2> gpurun_out/bench_${tag}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-per-config > gpurun_out/ncu_launch_${tag}.log 2>&1
for w in c5 t1g c2 c3 c4; do
  python tools/prof_run.py $w 5 > gpurun_out/prof_run_${tag}_$w.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -f -o gpurun_out/prof_${tag}_$w \
      python tools/prof_run.py $w 5 > gpurun_out/ncu_${tag}_$w.log 2>&1
done
ncu --set full --clock-control none -k regex:hist_kernel -s 4 -c 1 -f -o gpurun_out/prof_${tag}_hist python tools/hist_run.py t1g > gpurun_out/ncu_${tag}_hist.log 2>&1
ls -la gpurun_out/*${tag}*
