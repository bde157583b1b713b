set -x
mkdir -p gpurun_out
python scripts/dbg_k3.py > gpurun_out/r2_dbg_k3.log 2>&1; cat gpurun_out/r2_dbg_k3.log
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest3.log
tail -12 gpurun_out/r2_pytest3.log
python scripts/k2k3_case.py 262144 64 128 5 > gpurun_out/r2_k2k3_after2.json 2>&1 && cat gpurun_out/r2_k2k3_after2.json && \
ncu --set full --clock-control none --import-source on -k regex:'composite' -c 8 -o gpurun_out/r2_k3_after2 python scripts/k2k3_case.py 262144 64 128 1 > gpurun_out/r2_k3_ncu.log 2>&1
tail -3 gpurun_out/r2_k3_ncu.log
