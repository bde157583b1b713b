set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py tests/test_gpu_fullsize.py -q -p no:cacheprovider 2>&1 | tail -4
python scripts/k2k3_case.py 262144 64 128 5 > gpurun_out/r2_k2k3_after6.json; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_k2k3_after6.json'))
print({k:(v['ms'],v['frac_hbm']) for k,v in d.items() if isinstance(v,dict)})
P
python scripts/sweep_k2k3.py 262144 > gpurun_out/r2_sweep_k2k3_mlp.jsonl 2> gpurun_out/r2_sweep.err
python - <<'P'
import json
for l in open('gpurun_out/r2_sweep_k2k3_mlp.jsonl'):
    d=json.loads(l)
    print(d['rays'],d['Nc'],d['Nf'],' '.join(f"{k}={v['frac_hbm']:.2f}" for k,v in d.items() if isinstance(v,dict) and 'frac_hbm' in v))
P
