set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest8.log
tail -5 gpurun_out/r2_pytest8.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench8.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['extra']['repeat_ms_per_step'], d['extra']['render_800x800_frames_per_s'], d['extra'].get('dropin'))"
python bench.py --steps 2 --warmup 3 --no-cfg4 --no-dropin --no-cpu-baseline --no-frame > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dgrad|wgrad' -s 4 -c 4 -o gpurun_out/r2_bwd_kernels python bench.py --steps 2 --warmup 3 --no-cfg4 --no-dropin --no-cpu-baseline --no-frame > gpurun_out/ncu_bwd.log 2>&1
tail -2 gpurun_out/ncu_bwd.log
