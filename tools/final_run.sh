set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r01i_pytest.txt; cat gpurun_out/r01i_pytest.txt
timeout 300 python bench.py > gpurun_out/r01i_bench.json 2> gpurun_out/r01i_bench.err; tail -c 600 gpurun_out/r01i_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01i_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r01i_bench_reference.json
for w in cfg3_n60_2000_M4 n32_2000_M4 n16_2000_M4 n8_2000_M2 cfg1_n2_10_M2 n128_2000_M4 n256_1250_M4 cfg5_n32_500_S64_E128_M2; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r01i_bench_sweep.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/r01i_bench_sweep.jsonl'):
    d=json.loads(l); print(d['config']['workload'], round(d['value'],2), round(d['e2e']['value'],2), round(d['roofline'].get('whole_eval_frac', d['roofline']['frac']),3))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_r01i.log 2>&1; tail -2 gpurun_out/ncu_r01i.log | cut -c 1-200
