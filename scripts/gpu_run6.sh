set -x
mkdir -p gpurun_out
# 1. split-precision fp32 eval: parity tests + smoke + fp32 bench frame
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_tc.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2_pytest6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest6.log
tail -6 gpurun_out/r2_pytest6.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --mode fp32 --no-cpu-baseline --no-dropin --no-cfg4 --steps 5 > gpurun_out/r2_bench_fp32_split.json 2> gpurun_out/r2_bench_fp32_split.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_fp32_split.json') if l.startswith('{')][-1]); print('fp32 mode', d['ms_per_step'], d['extra'])"
# 2. inference forward speed: fp16+satfinite (shipped) vs fp16 plain vs timing noise
for rep in 1 2; do python scripts/perf_field.py bf16 8192 192 0; done
cd nerf_sandbox_b200/csrc && cp ../libnsb.so /tmp/libnsb_keep.so && nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DNSB_F16_SAT=0 -c -o field_tc.o field_tc.cu && nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ../libnsb.so engine.o sampler.o compositor.o field_fp32.o field_tc.o rays.o && cd ../..
for rep in 1 2; do python scripts/perf_field.py bf16 8192 192 0; done
cp /tmp/libnsb_keep.so nerf_sandbox_b200/libnsb.so
