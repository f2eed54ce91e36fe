set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t23_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t23_pytest.log
timeout 200 python scripts/bench_conv.py fwd 2>&1 | cut -c1-50
timeout 200 python scripts/bench_conv.py rows 2>&1 | cut -c1-50
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/t23_bench.json 2> gpurun_out/t23_bench.err; echo "bench rc=$?"
cp gpurun_out/bench_detail.json gpurun_out/t23_bench_detail.json
python scripts/show_detail.py 16
