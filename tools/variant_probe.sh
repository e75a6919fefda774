#!/bin/bash
# A/B of experimental library variants (build/variants/*.so) on the streaming step at N = 1e9.
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== variant $v"
  if [ "$v" = "base" ]; then unset PIC_LIB_PATH; else export PIC_LIB_PATH=build/variants/$v.so; fi
  python tools/stream_probe.py --deps ${DEPS:-split32} --tag _$v ${PROBE_ARGS}
done 2>&1 | tee gpurun_out/variant_probe.log
