#!/bin/bash
# GPU-box visit: parity tests + ncu --set full of one kernel inside the (calibrated) bench command.
# gpurun --timeout 1200 -- 'bash tools/gpu_check5.sh <tag> <kernel regex> <skip>'
TAG=${1:-run}; KRE=${2:-decode}; SKIP=${3:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP} -c 1 -f -o gpurun_out/${TAG}_${KRE} \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-train > gpurun_out/${TAG}_ncu_full.log 2>&1
echo done
