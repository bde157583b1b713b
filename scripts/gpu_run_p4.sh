set -x
N=${1:-4}
mkdir -p gpurun_out
NSB_TEST_NPROC=$N timeout 600 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider > gpurun_out/r2_p4_dist_${N}gpu.log 2>&1; echo "dist exit $?" >> gpurun_out/r2_p4_dist_${N}gpu.log
tail -12 gpurun_out/r2_p4_dist_${N}gpu.log
for tp in 1 0; do
NSB_TWO_PHASE=$tp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$tp bench.py --gpus $N --steps 40 --warmup 10 --no-frame --no-cfg4 --no-dropin --no-cpu-baseline > gpurun_out/r2_p4_bench_${N}gpu_tp$tp.json 2> gpurun_out/r2_p4_bench_${N}gpu_tp$tp.err; echo "bench exit $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_p4_bench_${N}gpu_tp$tp.json') if l.startswith('{')][-1]); print('two_phase=$tp', d['value'], d['ms_per_step'], d['extra']['repeat_ms_per_step'])"
done
python bench.py --steps 40 --warmup 10 --no-frame --no-cfg4 --no-dropin --no-cpu-baseline > gpurun_out/r2_p4_bench_1gpu.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_p4_bench_1gpu.json') if l.startswith('{')][-1]); print('single', d['value'], d['ms_per_step'], d['extra']['repeat_ms_per_step'])"
