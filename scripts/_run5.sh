set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t5_pytest.log
timeout 300 python scripts/bench_conv.py all 2>&1 | tee gpurun_out/t5_conv.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/t5_bench.json 2> gpurun_out/t5_bench.err; echo "bench rc=$?"
cp gpurun_out/bench_detail.json gpurun_out/t5_bench_detail.json
python scripts/show_detail.py 30
