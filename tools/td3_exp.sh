#!/bin/bash
# run the TD3 perf probe against experiment builds of the library (development aid)
cd "$(dirname "$0")/.."
C=residual-td3-robot-navigation_b200/csrc
cp $C/librtd3.so /tmp/base.so
for v in base NOCOMPUTE NOSTAGE; do
  if [ $v != base ]; then cp $C/librtd3_$v.so $C/librtd3.so; fi
  echo "== $v"; python tools/td3_prof.py 256 256 2 100 2>&1 | tail -1
  cp /tmp/base.so $C/librtd3.so
done
