#!/bin/bash
# dev helper: time prebuilt library variants gpurun_variants_<name>.so with the short bench (stage times only) and check parity
mkdir -p gpurun_out
for f in gpurun_variants_*.so; do
  name=${f#gpurun_variants_}; name=${name%.so}
  cp $f aeonflux_b200/csrc/libaeonflux_b200.so
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err || tail -3 gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/var_%s.json' % name).read().strip().splitlines()[-1])
    print(name, "value %.0f ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print(name, "failed", e)
PY
  if [ "$1" == "test" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_shapes or readme4_batch" 2>&1 | tail -2; fi
done
