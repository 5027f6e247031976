#!/bin/bash
# Multi-GPU measurement on an N-GPU box: bench.py under torchrun (the driver's launch), the reference arm, the one-process form.
# usage: bash tools/gpu_multi.sh <N> <tag> [extra bench flags]
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 "$@" \
    > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err || tail -20 gpurun_out/bench_${tag}_${N}gpu.err
python tools/bench_threads.py 5 > gpurun_out/bench_threads_${tag}_${N}gpu.json 2>> gpurun_out/bench_${tag}_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${tag}_${N}gpu.json').read().strip().splitlines()[-1])
print("N=%d value %.0f e2e %.0f streamed %.0f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["e2e"]["streamed_value"]))
s=d["secondary"]["stream_config5"]; print("stream", s["value"], s["wall_s"], s["mismatches"], s["input"])
print(open('gpurun_out/bench_threads_${tag}_${N}gpu.json').read())
PY
if [ "$N" == "2" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "multi or shard or stream" 2>&1 | tail -3; fi
