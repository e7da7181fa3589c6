mkdir -p gpurun_out
run() { env "$@" timeout 100 python tools/train_bench.py --steps 15 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys;b=json.loads(sys.stdin.read());print(sys.argv[1], round(b['value'],1))" "$*" >> gpurun_out/ab_bn_grid.txt; }
run ABCNET_BN_BLOCKS_PER_SM=16
run ABCNET_BN_BLOCKS_PER_SM=12
run ABCNET_BN_BLOCKS_PER_SM=16
run ABCNET_BN_BLOCKS_PER_SM=12
run ABCNET_BN_BLOCKS_PER_SM=8
cat gpurun_out/ab_bn_grid.txt
