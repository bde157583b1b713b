set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_final.log
tail -4 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; cat gpurun_out/r2_bench_reference_arm.json
timeout 600 python bench.py > gpurun_out/r2_bench_bf16.json 2> gpurun_out/r2_bench_bf16.err; echo "bench exit $?"; cat gpurun_out/r2_bench_bf16.json
timeout 600 python bench.py --mode fp32 --no-cpu-baseline --no-dropin > gpurun_out/r2_bench_fp32.json 2> gpurun_out/r2_bench_fp32.err; cat gpurun_out/r2_bench_fp32.json
python scripts/eval_chunk_sweep.py bf16 > gpurun_out/r2_eval_chunk_sweep.jsonl 2> gpurun_out/r2_eval_chunk_sweep.err; cat gpurun_out/r2_eval_chunk_sweep.jsonl
python scripts/sweep_k2k3.py 262144 > gpurun_out/r2_sweep_k2k3_mlp.jsonl 2> gpurun_out/r2_sweep.err; tail -2 gpurun_out/r2_sweep_k2k3_mlp.jsonl
python scripts/perf_bwd.py bf16 1024 192 > gpurun_out/r2_perf_bwd.json 2>&1; cat gpurun_out/r2_perf_bwd.json
python bench.py --steps 2 --warmup 3 --no-cfg4 --no-dropin --no-cpu-baseline --no-frame > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_train_step_launches.csv python bench.py --steps 2 --warmup 3 --no-cfg4 --no-dropin --no-cpu-baseline --no-frame > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'field_' -s 6 -c 6 -o gpurun_out/r2_field_kernels python scripts/perf_bwd.py bf16 1024 192 > gpurun_out/ncu_field.log 2>&1
tail -2 gpurun_out/ncu_field.log
