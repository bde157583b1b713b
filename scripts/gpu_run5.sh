set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest5.log
tail -8 gpurun_out/r2_pytest5.log
python scripts/k2k3_case.py 262144 64 128 5 > gpurun_out/r2_k2k3_after3.json 2>&1 && cat gpurun_out/r2_k2k3_after3.json && \
ncu --set full --clock-control none --import-source on -k regex:'stratified|resample' -c 6 -o gpurun_out/r2_k2_after3 python scripts/k2k3_case.py 262144 64 128 1 > gpurun_out/r2_k2_ncu.log 2>&1
tail -3 gpurun_out/r2_k2_ncu.log
