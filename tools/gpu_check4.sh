#!/bin/bash
# GPU-box visit: parity tests, bench line, training-step kernel breakdown, ncu --set full of one conv_igemm launch.
# gpurun --timeout 1200 -- 'bash tools/gpu_check4.sh <tag> <conv_igemm launch index to capture> [batch]'
TAG=${1:-run}; SKIP=${2:-33}; PB=${3:-256}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 1200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python tools/prof_train.py 64 > gpurun_out/${TAG}_prof_train.log 2>&1
echo "prof_train exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s ${SKIP} -c 1 -f -o gpurun_out/${TAG}_igemm_s${SKIP} \
    python tools/prof_forward.py ${PB} > gpurun_out/${TAG}_ncu_full.log 2>&1
echo done
