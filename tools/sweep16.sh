set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_shim.py tests/test_reference_run.py -m gpu -x -q > gpurun_out/s16_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s16_pytest.log
A="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $A > gpurun_out/s16_c4_1.json 2> gpurun_out/s16_err.log
python bench.py $A --dbuf 0 > gpurun_out/s16_c4_2.json 2>> gpurun_out/s16_err.log
python bench.py $A --cta-mode 2 > gpurun_out/s16_c4_3.json 2>> gpurun_out/s16_err.log
python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/s16_c5_1.json 2>> gpurun_out/s16_err.log
python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --cta-mode 2 > gpurun_out/s16_c5_2.json 2>> gpurun_out/s16_err.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s16_e2e.json 2>> gpurun_out/s16_err.log
tail -5 gpurun_out/s16_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s16_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], round(d['ms_per_step'],1), 'ms', 'e2e', (d.get('e2e') or {}).get('ms_per_step'), ((d.get('e2e') or {}).get('pinned') or {}).get('ms_per_step'))
PY
