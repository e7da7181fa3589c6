#!/bin/bash
# Round-end style visit: full parity suite, smoke(), both bench arms, ncu launch list of the bench command, ncu --set full
# captures of the dominant kernel and of one bandwidth-class kernel.   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh <tag>'
TAG=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv|decode|heads" -s 191 -c 96 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-train > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 37 -c 1 -f -o gpurun_out/${TAG}_heads_conv1 \
    python tools/prof_forward.py 256 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_c1 -c 1 -f -o gpurun_out/${TAG}_first_conv \
    python tools/prof_forward.py 256 >> gpurun_out/${TAG}_ncu_full.log 2>&1
echo done
