#!/bin/bash
# Round-2 closing visit on one B200: full parity suite, smoke(), both bench arms, and the per-kernel ncu table of the training-side
# bandwidth kernels (report kept on the box, only the text summary comes back).   gpurun --timeout 1200 -- 'bash tools/gpu_final2.sh <tag>'
TAG=${1:-r02_final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench.err
timeout 700 python bench.py > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
b = json.load(open("gpurun_out/${TAG}_bench.json"))
print({k: b[k] for k in ("value", "ms_per_step", "gpu_launches")}, b["e2e"]["value"], b["roofline"]["frac"], b.get("train", {}).get("ms_per_step"), b.get("train", {}).get("mem_gb"))
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread \
    --clock-control none -k "regex:bn_|loss_kernel|wgrad" -s 132 -c 132 -f -o /tmp/${TAG}_train \
    python tools/train_bench.py --mode eager --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_train.log 2>&1
{ echo "# ncu (time, DRAM bytes, DRAM throughput, registers) of the BatchNorm / loss / weight-gradient kernels of one eager training step at B = 64"
  echo "# (tools/train_bench.py --mode eager; window -s 132 -c 132 = one step's worth of these kernels). ncu's dram % is relative to 8.18 TB/s;"
  echo "# the measured copy peak is 6.545 TB/s. Cold-cache, serialised replays: read the GB/s column, not the absolute times."
  python tools/ncu_summary.py --group /tmp/${TAG}_train.ncu-rep; } > gpurun_out/${TAG}_train_kernels.summary.txt 2>&1
head -20 gpurun_out/${TAG}_train_kernels.summary.txt | cut -c1-160
echo done
