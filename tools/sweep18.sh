set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s18_pytest.log
A="--steps 3 --warmup 3 --no-cpu-baseline"
python bench.py $A > gpurun_out/s18_e2e_1.json 2> gpurun_out/s18_err.log
python bench.py $A --late-split 0 > gpurun_out/s18_e2e_2.json 2>> gpurun_out/s18_err.log
tail -5 gpurun_out/s18_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s18_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']; print(f.split('/')[-1], round(d['ms_per_step'],1), 'ms', 'e2e pageable', round(e['ms_per_step'],1), 'pinned', round(e['pinned']['ms_per_step'],1), 'device expect', round(e['device_expect_ms'],1), 'parity', d['parity']['ok'], d['parity']['rel_err_energy_e2e_vs_device'])
PY
