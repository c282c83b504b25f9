set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/s15_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s15_pytest.log
timeout 600 python tools/time_batch.py > gpurun_out/s15_batch.json 2> gpurun_out/s15_err.log; cat gpurun_out/s15_batch.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/s15_c4_1.json 2>> gpurun_out/s15_err.log
python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/s15_c5_1.json 2>> gpurun_out/s15_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s15_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], round(d['ms_per_step'],1), 'ms')
PY
