set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest9.log
tail -8 gpurun_out/r2_pytest9.log
timeout 600 python bench.py --no-cpu-baseline --no-cfg4 --no-frame > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench9.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value']); print(d['extra'].get('dropin')); print(d['extra'].get('dropin_reference_body'))"
