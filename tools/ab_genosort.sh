set -x
python -m pytest tests/test_gpu_models.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "geno or hier or replicate or fullsize" 2>&1 | tail -3
for v in "BB_GENO_SORT=0" "BB_GENO_SORT=1" "BB_GENO_SORT=1 BB_HZ_TRANSPOSE=1" "BB_GENO_SORT=0 BB_HZ_TRANSPOSE=1"; do
  echo "== cfg5 $v"; env $v QCFG=5 python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('step_us','p1_us','p2_us')})"
done
for v in "A=0" "BB_HZ_TRANSPOSE=1"; do
  echo "== cfg3 $v"; env $v QCFG=3 python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('step_us','p1_us','p2_us')})"
done
for v in "BB_GENO_SORT=0" "BB_GENO_SORT=1"; do
  echo "== cfg5/8 $v"; env $v QCFG=5 QSCALE=0.125 python tests/_quickbench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('step_us','p1_us','p2_us')})"
done
