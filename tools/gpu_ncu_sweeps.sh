# usage: bash tools/gpu_ncu_sweeps.sh TAG - ncu --set full of the sweep-side kernels of the default workload (raw pages as CSV)
set -x
TAG=${1:-r03}
mkdir -p gpurun_out
for k in k_sweep_fwd k_boundary_fwd k_magnus_adj; do
timeout 300 ncu --set full --clock-control none -k regex:^$k\|::$k -s 2 -c 1 -f -o /tmp/${TAG}_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$k.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_$k.log | cut -c 1-160
ncu -i /tmp/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${k}_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_ncu_${k}_raw.csv
done
