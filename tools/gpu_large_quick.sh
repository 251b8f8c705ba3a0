set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_baseline_sizes.py tests/test_gpu_units.py -m gpu -q --tb=short -k "large or n256 or n72" 2>&1 | tail -12
for w in n128_2000_M4 n256_1250_M4; do
for mode in ownlu cublaslu; do
  if [ $mode = cublaslu ]; then export QOCB_LARGE_CUBLAS_LU=1; else unset QOCB_LARGE_CUBLAS_LU; fi
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_$w.err | tail -1 > gpurun_out/${TAG}_bench_${w}_$mode.json; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${w}_$mode.json').read())
print('$w $mode value',d['value'],'stage_ms',{k:round(v,2) for k,v in d['stage_ms'].items()},'parity',d.get('parity'))
PY
done; done
unset QOCB_LARGE_CUBLAS_LU
timeout 600 python bench.py --workload n256_1250_M4 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_n256_1250_M4.json; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n256_1250_M4.json').read()); print('n256 full line: value',d['value'],'parity',d.get('parity'),'frac',d['roofline']['frac'],'whole',d['roofline']['whole_eval_frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_n256.csv python bench.py --workload n256_1250_M4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_n256.log 2>&1; python tools/summarize_launches.py gpurun_out/${TAG}_launches_n256.csv 2>/dev/null | sort -k6 -n -r | head -14
