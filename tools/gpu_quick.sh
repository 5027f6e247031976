#!/bin/bash
# quick GPU check used during development: short bench (stage times) + GPU parity tests
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])
    print("value %.0f  e2e %.0f  ms/step %.2f  pipeline_frac %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["pipeline_frac_of_imad_peak"]))
    print({k: round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print("bench failed", e)
PY
if [ "$1" != "notest" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
