set -x
mkdir -p gpurun_out
python scripts/tc_precision.py > gpurun_out/r2_tc_precision.json 2> gpurun_out/r2_tc_precision.err; cat gpurun_out/r2_tc_precision.json; tail -3 gpurun_out/r2_tc_precision.err
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest4.log
tail -12 gpurun_out/r2_pytest4.log
timeout 300 python bench.py --no-cfg4 --no-dropin --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; cat gpurun_out/r2_bench4.json
