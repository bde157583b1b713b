set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split_gemm.py tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_graph.py tests/test_gpu_fullsize.py -q -p no:cacheprovider 2>&1 | tail -6
python scripts/perf_split_gemm.py 2>&1 | tee gpurun_out/r2_split_gemm_perf.txt
for cfg in "1 3" "0 3" "1 2"; do set -- $cfg; NSB_SPLIT_WGRAD_WIDE=$1 NSB_SPLIT_TERMS=$2 python scripts/perf_step.py 1024 fp32 2>&1 | tail -1 | sed "s/^/[wide=$1 nt=$2] /"; done | tee gpurun_out/r2_fp32_step.txt
