set -x
mkdir -p gpurun_out
python scripts/dbg_split.py 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest7.log
tail -6 gpurun_out/r2_pytest7.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --mode fp32 --no-cpu-baseline --no-dropin --no-cfg4 --steps 5 > gpurun_out/r2_bench_fp32_split.json 2> gpurun_out/r2_bench_fp32_split.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_fp32_split.json') if l.startswith('{')][-1]); print('fp32 mode', d['ms_per_step'], d['extra'])"
