set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t18_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t18_pytest.log
timeout 200 python scripts/bench_conv.py rows 2>&1 | cut -c1-50
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/t18_bench.json 2> gpurun_out/t18_bench.err; echo "bench rc=$?"
cp gpurun_out/bench_detail.json gpurun_out/t18_bench_detail.json
python scripts/show_detail.py 70 | grep "64>128k3\|96>128k3\|128>96k3\|sum of\|64>192"
