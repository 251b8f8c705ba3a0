set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r02a_pytest.txt; tail -5 gpurun_out/r02a_pytest.txt
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.json
