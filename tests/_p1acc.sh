#!/bin/bash
# development helper (GPU box): pass-1 accumulator budget (shared memory <-> resident CTAs) on the hierarchical configurations
cd "$(dirname "$0")/.."
for c in 3 5 4; do
  for kb in 44 32 20 12; do
    echo "cfg$c BB_P1_ACC_KB=$kb"
    BB_P1_ACC_KB=$kb QCFG=$c QN=200 timeout 300 python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(d[k],2) for k in ('step_us','p1_us','p2_us')})"
  done
done
