#!/bin/bash
# Round measurement pass on one B200.  Usage: bash tools/gpu_profile.sh <tag> [bench|ncu|all]
#   bench: full bench line (+ reference arm), S16 pass-size sweep, ncu launch list of the bench command
#   ncu:   `ncu --set full` captures of the ladder kernels at README-4, S16 and issuance; only their summaries (tools/ncu_summary.py)
#          and source-level instruction mixes leave the box (gpurun_out/ is capped at 64 MiB)
tag=${1:-v}; what=${2:-all}
mkdir -p gpurun_out
if [ "$what" == "bench" ] || [ "$what" == "all" ]; then
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_$tag.err
python tools/s16_sweep.py > gpurun_out/s16_sweep_$tag.json 2>> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-stream > gpurun_out/b_ncu_$tag.log 2>&1
tail -c 600 gpurun_out/bench_$tag.json | head -c 600; echo
cat gpurun_out/bench_${tag}_ref.json | head -c 400; echo
fi
if [ "$what" == "ncu" ] || [ "$what" == "all" ]; then
cap() {   # cap <name> <kernel regex> <skip> <count> <command...>
    name=$1; regex=$2; skip=$3; count=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_full_${name}_$tag.log 2>&1
    python tools/ncu_summary.py /tmp/prof_$name.ncu-rep > gpurun_out/ncu_full_${name}_${tag}_summary.csv
    ncu -i /tmp/prof_$name.ncu-rep --page source --csv --print-source sass > /tmp/src_$name.csv 2>/dev/null
    python tools/sass_mix.py /tmp/src_$name.csv > gpurun_out/sass_mix_${name}_$tag.txt 2>&1
    rm -f /tmp/src_$name.csv
}
cap readme4 'k_ladders|k_points' 6 2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-stream
cap s16 'k_ladders|k_points' 2 2 python tools/profile_secondary.py s16
cap issue 'k_msm_ct|k_ladders|k_points' 4 4 python tools/profile_secondary.py issue
ls -la /tmp/prof_*.ncu-rep
fi
