set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_baseline_sizes.py tests/test_gpu_units.py -m gpu -q --tb=short -x -k "large or n256 or n72" 2>&1 | tail -8
for w in n128_2000_M4 n256_1250_M4; do
for mode in tma cpasync cublas; do
  if [ $mode = cpasync ]; then export QOCB_NO_TMA=1; else unset QOCB_NO_TMA; fi
  if [ $mode = cublas ]; then export QOCB_LARGE_CUBLAS=1; else unset QOCB_LARGE_CUBLAS; fi
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_$w.err | tail -1 > gpurun_out/${TAG}_bench_${w}_$mode.json; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${w}_$mode.json').read())
print('$w $mode value',d['value'],'stage_ms',{k:round(v,2) for k,v in d['stage_ms'].items()},'parity',d.get('parity'))
PY
done; done
