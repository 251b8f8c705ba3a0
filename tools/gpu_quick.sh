# usage: bash tools/gpu_quick.sh TAG ["pytest files"] - smoke, a subset of GPU tests, default bench line, phase profile
set -x
TAG=${1:-r02x}
SEL=${2:-"tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_baseline_sizes.py"}
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest $SEL -m gpu -q --tb=short -x 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.txt; tail -8 gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['kernel'],'stage_ms',d['stage_ms'],'parity',d.get('parity'))
PY
timeout 300 python tools/phase_profile.py > gpurun_out/${TAG}_phases.txt 2>&1; cat gpurun_out/${TAG}_phases.txt
