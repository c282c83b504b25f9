set -x
timeout 1500 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s11_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/s11_c4_1.json 2> gpurun_out/s11_err.log || tail -5 gpurun_out/s11_err.log
python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/s11_c5_1.json 2>> gpurun_out/s11_err.log || tail -5 gpurun_out/s11_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s11_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], round(d['ms_per_step'],1), 'ms')
PY
