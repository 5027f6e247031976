#!/bin/bash
# compute-sanitizer over every batch operation (small sizes).  usage: bash tools/gpu_sanitize.sh <tag>
tag=${1:-r}
out=gpurun_out/compute_sanitizer_$tag.txt
mkdir -p gpurun_out; : > $out
for tool in memcheck racecheck synccheck; do
    echo "== compute-sanitizer --tool $tool python tests/tools/sanitize_small.py" >> $out
    timeout 1500 compute-sanitizer --tool $tool python tests/tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -12 >> $out
done
tail -40 $out
