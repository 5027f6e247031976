#!/bin/bash
# dev helper: like gpu_variants.sh, plus the DRAM traffic of one k_ladders launch per variant (ncu, single pass of three metrics)
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
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_ladders -s 4 -c 1 --csv \
      --log-file gpurun_out/traffic_$name.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/traffic_$name.csv | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}'
done
