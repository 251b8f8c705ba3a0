# usage: bash tools/gpu_large.sh TAG - large-dimension tests + n128 / n256 bench lines + launch list
set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_baseline_sizes.py tests/test_gpu_units.py -m gpu -q --tb=short -x 2>&1 | tail -15
for w in n128_2000_M4 n256_1250_M4; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 2>gpurun_out/${TAG}_$w.err | tail -1 > gpurun_out/${TAG}_bench_$w.json; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_$w.json').read())
print('$w','value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['kernel'],'whole',d['roofline']['whole_eval_frac'],'stage_ms',d['stage_ms'],'parity',d.get('parity'))
PY
done
QOCB_LARGE_CUBLAS=1 timeout 600 python bench.py --workload n256_1250_M4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_n256_cublasgemm.json; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n256_cublasgemm.json').read()); print('n256 with cuBLAS ZGEMM: value',d['value'],'stage_ms',d['stage_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_n256.csv python bench.py --workload n256_1250_M4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_n256.log 2>&1; python tools/summarize_launches.py gpurun_out/${TAG}_launches_n256.csv 2>/dev/null | head -30
