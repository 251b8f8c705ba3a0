# usage: bash tools/gpu_expm.sh TAG - expm parity tests + the batched-expm bench workloads into gpurun_out/TAG_expm.jsonl
set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -m gpu -q --tb=short -x -k "expm" 2>&1 | tail -15
rm -f gpurun_out/${TAG}_expm.jsonl
for n in 2 4 8 16 32 64; do timeout 300 python bench.py --workload expm_batched_n$n --steps 10 --warmup 3 2>gpurun_out/${TAG}_expm.err | tail -1 >> gpurun_out/${TAG}_expm.jsonl; done
python - <<PY
import json
for l in open('gpurun_out/${TAG}_expm.jsonl'):
    d=json.loads(l); r=d['roofline']
    print(d['config']['workload'], 'mat/s %.3g'%d['matrices_per_s'], 'GFLOP/s %.0f'%d['value'], r['bound'], 'frac %.3f'%r['frac'], 'e2e mat/s %.3g'%d['e2e']['matrices_per_s'], 'cpu mat/s %.3g'%d['cpu_baseline']['matrices_per_s'], 'parity %.1e'%d['parity']['rel_err'])
PY
