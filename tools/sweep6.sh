set -x
timeout 1500 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s6_pytest.log
i=0
for args in "" "--low-bits 3" "--reg-bits 4" "--reg-bits 4 --low-bits 3" "--low-bits 3 --cta-mode 2" "--low-bits 3 --tile-bits 13"; do
  i=$((i+1))
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity $args > gpurun_out/s6_c4_$i.json 2> gpurun_out/s6_err.log || tail -5 gpurun_out/s6_err.log
done
i=0
for args in "" "--cta-mode 1" "--low-bits 3"; do
  i=$((i+1))
  python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity $args > gpurun_out/s6_c5_$i.json 2>> gpurun_out/s6_err.log || tail -5 gpurun_out/s6_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s6_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], 'rb', d['config']['reg_bits'], 'tile', d['config']['tile_bits'], round(d['ms_per_step'],1), 'ms  energy', d['energy'])
PY
