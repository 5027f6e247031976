mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_ladders' -s 3 -c 1 -f -o /tmp/prof_lad python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-stream > gpurun_out/cap_src.log 2>&1
ncu -i /tmp/prof_lad.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/src_k_ladders.csv.gz
ls -la gpurun_out/src_k_ladders.csv.gz /tmp/prof_lad.ncu-rep
