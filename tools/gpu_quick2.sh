# usage: bash tools/gpu_quick2.sh TAG - a few GPU tests, default bench line, cfg5 small line, boundary phase profile
set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_units.py -m gpu -q --tb=short -x 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload cfg5_n32_500_S64_E128_M2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err
python - <<PY
import json
for f in ('bench','cfg5'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'stage_ms',d['stage_ms'],'parity',d.get('parity'))
    except Exception as e: print(f,'failed',e)
PY
timeout 300 python tools/phase_profile.py 2>&1 | grep "us/step\|slices sampled"
