# usage: bash tools/gpu_sweep.sh TAG - full GPU tests, default bench, the sweep of workloads
set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.txt; tail -4 gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
rm -f gpurun_out/${TAG}_bench_sweep.jsonl
for w in cfg3_n60_2000_M4 n32_2000_M4 n16_2000_M4 n8_2000_M2 cfg1_n2_10_M2 n128_2000_M4 n256_1250_M4 cfg5_n32_500_S64_E128_M2; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/${TAG}_bench_sweep.jsonl; done
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('default', d['value'], d['e2e']['value'], d['roofline']['frac'], d['stage_ms'], d.get('parity'))
for l in open('gpurun_out/${TAG}_bench_sweep.jsonl'):
    d=json.loads(l); print(d['config']['workload'], round(d['value'],2), round(d['e2e']['value'],2), round(d['roofline']['frac'],3), d['stage_ms'], d.get('parity'))
PY
