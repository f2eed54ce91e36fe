#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one full capture of the top kernels.
# usage: scripts/gpu_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
set -u
tag=${1:-r01}
mkdir -p gpurun_out
timeout ${PYTEST_TIMEOUT:-600} python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cp gpurun_out/bench_detail.json gpurun_out/${tag}_bench_detail.json 2>/dev/null
CMD="timeout ${NCU_TIMEOUT:-420} python bench.py --steps 1 --warmup 3 --value-only"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-conv_tc}" -s ${NCU_SKIP:-400} -c ${NCU_COUNT:-6} -f -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
# DRAM traffic of every conv-family launch of the run (3 warm-up + 1 timed step; one cheap counter pass)
$CMD > gpurun_out/${tag}_plain3.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"conv_rows|conv_tc|wgrad_rows|conv_simt|colsum" -c 4000 --csv --log-file gpurun_out/${tag}_traffic.csv \
    $CMD > gpurun_out/${tag}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
if [ -n "${NCU_STEP_SKIP:-}" ]; then
  # light counters for every launch of one whole step (no source, few sections)
  $CMD > gpurun_out/${tag}_plain3.log 2>&1 &&
  ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none \
      -s ${NCU_STEP_SKIP} -c ${NCU_STEP_COUNT:-400} -f -o gpurun_out/${tag}_step $CMD > gpurun_out/${tag}_ncu_step.log 2>&1
  echo "ncu step rc=$?"
fi
ls -la gpurun_out | tail -20
