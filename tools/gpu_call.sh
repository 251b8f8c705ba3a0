# usage: bash tools/gpu_call.sh TAG [profile]  - smoke, GPU tests, default bench line (and the phase profile) into gpurun_out/TAG_*
set -x
TAG=${1:-r02x}
mkdir -p gpurun_out
if ! timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; then
  tail -5 gpurun_out/${TAG}_smoke.txt
  echo "SMOKE FAILED - retrying without TMA, then without the Magnus pre-pass"
  QOCB_NO_TMA=1 timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
  QOCB_NO_TMA=1 QOCB_NO_PREMAGNUS=1 timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
  exit 1
fi
tail -2 gpurun_out/${TAG}_smoke.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.txt; tail -5 gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['kernel'],'stage_ms',d['stage_ms'],'parity',d.get('parity'))
PY
if [ "$2" = "profile" ]; then timeout 300 python tools/phase_profile.py > gpurun_out/${TAG}_phases.txt 2>&1; cat gpurun_out/${TAG}_phases.txt; fi
if [ "$3" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forward -s 2 -c 1 -f -o gpurun_out/${TAG}_kforward python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log | cut -c 1-300
fi
