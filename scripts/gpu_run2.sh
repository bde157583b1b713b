set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
python scripts/k2k3_case.py 262144 64 128 5 > gpurun_out/r2_k2k3_after1.json 2>&1 && cat gpurun_out/r2_k2k3_after1.json && \
ncu --set full --clock-control none --import-source on -k regex:'composite' -c 8 -o gpurun_out/r2_k3_after1 python scripts/k2k3_case.py 262144 64 128 1 > gpurun_out/r2_k3_ncu.log 2>&1
tail -3 gpurun_out/r2_k3_ncu.log
timeout 300 python bench.py --no-cfg4 --no-dropin --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; cat gpurun_out/r2_bench2.json
