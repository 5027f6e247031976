#!/bin/bash
# dev helper: time prebuilt library variants gpurun_variants_ct*.so on the prover path (tools/time_prover.py: device-resident
# Issuer::issue, its output verified and digested -- every variant must print the same digest)
mkdir -p gpurun_out
cp aeonflux_b200/csrc/libaeonflux_b200.so /tmp/lib_keep.so
for f in gpurun_variants_ct*.so; do
  name=${f#gpurun_variants_}; name=${name%.so}
  cp $f aeonflux_b200/csrc/libaeonflux_b200.so
  python tools/time_prover.py > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err || tail -3 gpurun_out/var_$name.err
  echo "$name $(cat gpurun_out/var_$name.json)"
done
cp /tmp/lib_keep.so aeonflux_b200/csrc/libaeonflux_b200.so
