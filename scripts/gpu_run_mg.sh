set -x
N=${1:-2}
mkdir -p gpurun_out
NSB_TEST_NPROC=$N timeout 900 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider > gpurun_out/r2_dist_test_${N}gpu.log 2>&1; echo "dist exit $?" >> gpurun_out/r2_dist_test_${N}gpu.log
tail -15 gpurun_out/r2_dist_test_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench exit $?"
cat gpurun_out/r2_bench_${N}gpu.json; tail -5 gpurun_out/r2_bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29602 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_${N}gpu.json 2> gpurun_out/r2_bench_ref_${N}gpu.err; cat gpurun_out/r2_bench_ref_${N}gpu.json
