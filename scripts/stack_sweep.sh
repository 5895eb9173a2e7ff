#!/bin/bash
# usage: scripts/stack_sweep.sh "VAR=1 VAR2=2" "VAR=3" ...   (each argument = one environment; run on the GPU box)
python scripts/stack_time.py 3000 40 baseline
for e in "$@"; do
  env $e python scripts/stack_time.py 3000 40 "$e"
done
