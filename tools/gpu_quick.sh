#!/bin/bash
# quick GPU check used during development: short bench (stage times) + GPU parity tests.  usage: gpu_quick.sh [notest] [bench flags...]
mkdir -p gpurun_out
notest=0; if [ "$1" == "notest" ]; then notest=1; shift; fi
python bench.py --steps 3 --warmup 3 "$@" > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -20 gpurun_out/quick.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])
    print("value %.0f  e2e %.0f  ms/step %.2f  pipeline_frac %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["pipeline_frac_of_imad_peak"]))
    print({k: round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()})
    for k,v in (d.get("secondary") or {}).items():
        print(k, "%.0f %s" % (v["value"], v["unit"]), "e2e %.0f" % v["e2e"]["value"] if "e2e" in v else "", "frac %.3f" % v["frac"] if "frac" in v else "",
              {kk: round(vv.get("frac_of_imad_peak", vv.get("frac_of_alu_peak", 0)), 3) for kk, vv in (v.get("kernels") or {}).items()}, (v.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("bench failed", e)
PY
if [ $notest == 0 ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6; fi
