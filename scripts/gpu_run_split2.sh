set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split_gemm.py tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_gpu_graph.py tests/test_gpu_fullsize.py -q -p no:cacheprovider 2>&1 | tail -12
python scripts/perf_split_gemm.py 2>&1 | tee gpurun_out/r2_split_gemm_perf.txt
for cfg in "tc 3" "tc 2"; do set -- $cfg; NSB_FP32_GEMM=$1 NSB_SPLIT_TERMS=$2 python scripts/perf_step.py 1024 fp32 2>&1 | tail -1 | sed "s/^/[$1 nt=$2] /"; done | tee gpurun_out/r2_fp32_step.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_fp32_step_launches.csv python scripts/perf_step.py 1024 fp32 > gpurun_out/ncu_fp32.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_fp32_step_launches.csv 2>&1 | head -14
