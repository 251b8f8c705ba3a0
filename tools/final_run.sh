# usage: bash tools/final_run.sh TAG  - everything the round's profiles/ are built from (one GPU)
set -x
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; tail -c 400 gpurun_out/${TAG}_bench_reference.json
rm -f gpurun_out/${TAG}_bench_sweep.jsonl
for w in cfg3_n60_2000_M4 n32_2000_M4 n16_2000_M4 n8_2000_M2 cfg1_n2_10_M2 n128_2000_M4 n256_1250_M4 cfg5_n32_500_S64_E128_M2; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/${TAG}_bench_sweep.jsonl; done
timeout 300 python bench.py --workload cfg2_lindblad_n2 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_lindblad.json
python - <<PY
import json
for l in open('gpurun_out/${TAG}_bench_sweep.jsonl'):
    d=json.loads(l); print(d['config']['workload'], round(d['value'],2), round(d['e2e']['value'],2), round(d['roofline']['frac'],3), d.get('parity'))
PY
# ncu --set full of the two dominant kernels: the reports stay on the box (gpurun_out/ is limited to 64 MiB), their raw and
# source pages come back as CSV
for k in k_forward k_backward; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/${TAG}_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$k.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_$k.log | cut -c 1-200
ncu -i /tmp/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${k}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${TAG}_ncu_${k}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_ncu_${k}_*.csv
done
# launch list of the default bench command (per-launch durations; cold cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1; tail -2 gpurun_out/${TAG}_launches.log | cut -c 1-200
timeout 300 python tools/phase_profile.py > gpurun_out/${TAG}_phases.txt 2>&1; tail -45 gpurun_out/${TAG}_phases.txt
