set -x
timeout 1500 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_eigensolver.py -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s7_pytest.log
i=0
for args in "" "--min-bundle 3" "--min-bundle 4" "--min-bundle 255"; do
  i=$((i+1))
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity $args > gpurun_out/s7_c4_$i.json 2> gpurun_out/s7_err.log || tail -5 gpurun_out/s7_err.log
  python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity $args > gpurun_out/s7_c5_$i.json 2>> gpurun_out/s7_err.log || tail -5 gpurun_out/s7_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s7_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], 'sets', d['config']['partner_load_sets'], round(d['ms_per_step'],1), 'ms  energy', d['energy'])
PY
