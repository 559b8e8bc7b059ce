#!/bin/bash
# GPU box, one gpurun call: the ncu evidence of round 2 (each command first runs plainly and must exit 0)
#   1. launch list of a short bench.py run            -> gpurun_out/r2_ncu_launches.csv
#   2. --set full of the step kernel (cfg2, K = 8)     -> gpurun_out/prof_stepk.ncu-rep
#   3. --set full of the hierarchical pass 1 / pass 2  -> gpurun_out/prof_cfg3.ncu-rep
cd "$(dirname "$0")/.."
B="python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 20"
$B > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
export BB_PERSIST=0 QK=8 QSTEPS=8
QCFG=2 python tests/_ncu_target.py > gpurun_out/ncu_plain_stepk.log 2>&1 &&
QCFG=2 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/prof_stepk -f python tests/_ncu_target.py > gpurun_out/ncu_stepk.log 2>&1
echo "stepk rc=$?"
QCFG=3 python tests/_ncu_target.py > gpurun_out/ncu_plain_cfg3.log 2>&1 &&
QCFG=3 ncu --set full --clock-control none --import-source on -k regex:"pass[12]_kernel" -s 8 -c 2 -o gpurun_out/prof_cfg3 -f python tests/_ncu_target.py > gpurun_out/ncu_cfg3.log 2>&1
echo "cfg3 rc=$?"
ls -la gpurun_out/*.ncu-rep
