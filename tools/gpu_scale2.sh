set -x
bash tools/gpu_scale.sh r02r cfg4_n256_10000_M4 "1 8" --steps 3 --warmup 3 --no-cpu-baseline
bash tools/gpu_scale.sh r02r cfg5_n32_500_S64_E1024_M2 "1 8" --steps 3 --warmup 3 --no-cpu-baseline
