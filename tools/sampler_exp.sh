#!/bin/bash
cd "$(dirname "$0")/.."
C=residual-td3-robot-navigation_b200/csrc
cp $C/librtd3.so /tmp/base.so
for v in base s128 s64; do
  if [ $v != base ]; then cp $C/librtd3_$v.so $C/librtd3.so; fi
  echo "== $v"; python tools/sampler_prof.py 2>&1 | tail -2; python -m pytest tests/test_td3_gpu.py -q -k "sample" 2>&1 | tail -1
  cp /tmp/base.so $C/librtd3.so
done
