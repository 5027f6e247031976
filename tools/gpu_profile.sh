#!/bin/bash
# Round measurement pass on one B200: full bench line (+ reference arm), ncu launch list of the same command, and one
# `ncu --set full` capture of the ladder kernels.  Usage: bash tools/gpu_profile.sh <tag>
tag=${1:-v}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ladders|k_points' -s 6 -c 2 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full_$tag.log 2>&1
tail -c 600 gpurun_out/bench_$tag.json | head -c 600; echo
cat gpurun_out/bench_${tag}_ref.json | head -c 400; echo
