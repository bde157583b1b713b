set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest10.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest10.log
tail -6 gpurun_out/r2_pytest10.log
python scripts/k2k3_case.py 262144 64 128 5 > gpurun_out/r2_k2k3_after4.json 2>&1 && cat gpurun_out/r2_k2k3_after4.json && \
ncu --set full --clock-control none --import-source on -k regex:'resample' -c 4 -o gpurun_out/r2_k2_after4 python scripts/k2k3_case.py 262144 64 128 1 > gpurun_out/r2_k2_ncu.log 2>&1
tail -2 gpurun_out/r2_k2_ncu.log
