set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest1.log
tail -30 gpurun_out/r2_pytest1.log
timeout 600 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench exit $?"
cat gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
python scripts/k2k3_case.py 262144 64 128 5 > gpurun_out/r2_k2k3_before.json 2>&1 && cat gpurun_out/r2_k2k3_before.json && \
ncu --set full --clock-control none --import-source on -k regex:'stratified|resample|composite' -c 14 -o gpurun_out/r2_k2k3_before python scripts/k2k3_case.py 262144 64 128 1 > gpurun_out/r2_k2k3_ncu.log 2>&1
tail -3 gpurun_out/r2_k2k3_ncu.log
