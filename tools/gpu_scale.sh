# usage: bash tools/gpu_scale.sh TAG WORKLOAD "N..." [extra bench args]  - torchrun bench.py at each N into gpurun_out/TAG_scale_<workload>_gpusN.json
set -x
TAG=${1:-r02x}; WL=${2:-n64_2000_M4}; NS=${3:-"2 4 8"}; shift 3
mkdir -p gpurun_out
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --workload $WL "$@" > gpurun_out/${TAG}_scale_${WL}_gpus1.json 2> gpurun_out/${TAG}_scale_${WL}_gpus1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --workload $WL "$@" > gpurun_out/${TAG}_scale_${WL}_gpus$N.json 2> gpurun_out/${TAG}_scale_${WL}_gpus$N.err
  fi
  tail -c 1800 gpurun_out/${TAG}_scale_${WL}_gpus$N.json; tail -3 gpurun_out/${TAG}_scale_${WL}_gpus$N.err
done
