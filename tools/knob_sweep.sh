#!/bin/bash
# development: pair-sum rate against particle count for launch-plan knob settings (run on the GPU box)
cd "$(dirname "$0")/.."
for n in 5000 20000 30000 50000 70000 100000 200000 400000; do
  reps=20; [ $n -ge 300000 ] && reps=3
  for fb in 0 1000000; do
    O2D_PP_FINE_BELOW=$fb python tools/quick_rate.py $n $reps rm:1:0,v2:1:1,rm:2:0 2>&1 | grep -v "^$"
  done
done
