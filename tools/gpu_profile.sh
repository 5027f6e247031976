#!/bin/bash
# Round measurement pass on one B200: full bench line (+ reference arm), ncu launch list of the same command, and `ncu --set full`
# captures of the ladder kernels at README-4 and at S16.  Usage: bash tools/gpu_profile.sh <tag>
tag=${1:-v}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_$tag.err
python tools/s16_sweep.py > gpurun_out/s16_sweep_$tag.json 2>> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-stream > gpurun_out/b_ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ladders|k_points' -s 6 -c 2 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-stream > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ladders|k_points' -s 2 -c 2 -f -o gpurun_out/prof_s16_$tag \
    python tools/profile_secondary.py s16 > gpurun_out/ncu_full_s16_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_msm_ct|k_ladders' -s 2 -c 2 -f -o gpurun_out/prof_issue_$tag \
    python tools/profile_secondary.py issue > gpurun_out/ncu_full_issue_$tag.log 2>&1
tail -c 600 gpurun_out/bench_$tag.json | head -c 600; echo
cat gpurun_out/bench_${tag}_ref.json | head -c 400; echo
