# usage: bash tools/gpu_some.sh "pytest args"
set -x
timeout 1200 python -m pytest $1 -m gpu -q --tb=short 2>&1 | tail -25
