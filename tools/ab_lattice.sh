python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 3 4 5; do
  echo "== cfg$c"; QCFG=$c python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('step_us','p1_us','p2_us')})"
done
echo "== cfg2 K=1"; QCFG=2 QK=1 python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('step_us','p1_us','p2_us')})"
echo "== cfg2 truncated"; QCFG=2 QOPT=truncated python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('step_us','p1_us','p2_us')})"
echo "== cfg2/8"; QCFG=2 QSCALE=0.125 python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('step_us','p1_us','p2_us')})"
