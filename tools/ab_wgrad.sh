# A/B of the two tap-folded weight-gradient variants on the whole training step (one box, alternating runs)
mkdir -p gpurun_out
run() { env "$@" timeout 120 python tools/train_bench.py --steps 15 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys;b=json.loads(sys.stdin.read());print(sys.argv[1], round(b['value'],1))" "$*" >> gpurun_out/ab_wgrad.txt; }
for rep in 1 2; do
run ABCNET_X=0
run ABCNET_WGRAD_ROWBOX=1
done
cat gpurun_out/ab_wgrad.txt
