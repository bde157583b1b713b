set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_final.log
tail -5 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 600 gpurun_out/r2_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r2_bench_bf16.json 2> gpurun_out/r2_bench_bf16.err; echo "bench exit $?"; cat gpurun_out/r2_bench_bf16.json
