set -x
mkdir -p gpurun_out
for nt in 3 2; do NSB_SPLIT_TERMS=$nt python scripts/dbg_split_gemm.py 2>&1 | sed "s/^/[nt=$nt] /"; done | tee gpurun_out/r2_split_anatomy.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_fp32_step_launches.csv python scripts/perf_step.py 1024 fp32 > gpurun_out/ncu_fp32.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_fp32_step_launches.csv 2>&1 | head -40
