set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t4_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/t4_bench.json 2> gpurun_out/t4_bench.err; echo "bench rc=$?"
cp gpurun_out/bench_detail.json gpurun_out/t4_bench_detail.json
python scripts/show_detail.py 30
timeout 200 python scripts/bench_fe.py motion 2>&1 | tee gpurun_out/t4_motion.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"corr_grad|warp_bwd" -s 6 -c 3 -f -o gpurun_out/t4_corr python scripts/bench_fe.py motion > gpurun_out/t4_ncu.log 2>&1; echo "ncu rc=$?"
