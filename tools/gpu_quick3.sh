# usage: bash tools/gpu_quick3.sh TAG - GPU tests touching the sweeps, default bench with and without the three-level scheme
set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_units.py tests/test_gpu_baseline_sizes.py tests/test_gpu_sharded.py tests/test_gpu_timedep.py -m gpu -q --tb=short -x 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
QOCB_NO_THREE_LEVEL=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
timeout 600 python bench.py --workload n32_2000_M4 --no-cpu-baseline > gpurun_out/${TAG}_n32.json 2> gpurun_out/${TAG}_n32.err
python - <<PY
import json
for f in ('bench','bench2','n32'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'],'stage_ms',d['stage_ms'],'parity',d.get('parity'))
    except Exception as e: print(f,'failed',e)
PY
