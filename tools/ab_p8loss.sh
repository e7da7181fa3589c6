mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_ops_gpu.py tests/test_train_gpu.py tests/test_host_paths_gpu.py -x -q 2>&1 | tail -12 > gpurun_out/r02_p8loss_tests.log
tail -5 gpurun_out/r02_p8loss_tests.log
run() { env "$@" timeout 120 python tools/train_bench.py --steps 15 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys;b=json.loads(sys.stdin.read());print(sys.argv[1], round(b['value'],1), b.get('mem_gb'))" "$*" >> gpurun_out/ab_p8loss.txt; }
for rep in 1 2; do
run ABCNET_LOSS_FP32=1
run ABCNET_X=0
done
cat gpurun_out/ab_p8loss.txt
