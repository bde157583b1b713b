set -x
mkdir -p gpurun_out
NSB_SPLIT_TERMS=3 python scripts/dbg_split_gemm.py 2>&1 | sed "s/^/[nt=3] /" | tee gpurun_out/r2_split_anatomy.txt
timeout 900 python -m pytest tests/test_gpu_split_gemm.py tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_graph.py tests/test_gpu_fullsize.py -q -p no:cacheprovider > gpurun_out/r2_pytest_split.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_split.log
grep -n "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/r2_pytest_split.log | head -20
for cfg in "tc 3" "tc 2"; do set -- $cfg; NSB_FP32_GEMM=$1 NSB_SPLIT_TERMS=$2 python scripts/perf_step.py 1024 fp32 2>&1 | tail -1 | sed "s/^/[$1 nt=$2] /"; done | tee gpurun_out/r2_fp32_step.txt
