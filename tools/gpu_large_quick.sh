set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_baseline_sizes.py -m gpu -q --tb=short -x -k "large or n256" 2>&1 | tail -8
timeout 600 python bench.py --workload n256_1250_M4 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_n256.err | tail -1 > gpurun_out/${TAG}_bench_n256_1250_M4.json; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n256_1250_M4.json').read())
print('value',d['value'],'e2e',d['e2e']['value'],'stage_ms',d['stage_ms'],'parity',d.get('parity'))
PY
