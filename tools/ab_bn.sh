mkdir -p gpurun_out
run() { env "$@" timeout 120 python tools/train_bench.py --steps 15 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys;b=json.loads(sys.stdin.read());print(sys.argv[1], round(b['value'],1))" "$*" >> gpurun_out/ab_bn.txt; }
for rep in 1 2; do
run ABCNET_BN_MINB=32
run ABCNET_BN_MINB=33
run ABCNET_BN_MINB=34
run ABCNET_BN_MINB=44
run ABCNET_BN_MINB=33 ABCNET_BN_BLOCKS_PER_SM=12
run ABCNET_BN_MINB=33 ABCNET_BN_BLOCKS_PER_SM=24
run ABCNET_BN_MINB=34 ABCNET_BN_BLOCKS_PER_SM=24
done
cat gpurun_out/ab_bn.txt
