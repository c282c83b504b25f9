set -x
i=0
for args in "--tile-bits 13" "--tile-bits 13 --low-bits 4" "--low-bits 2" "--cta-mode 2"; do
  i=$((i+1))
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity $args > gpurun_out/s10_c4_$i.json 2> gpurun_out/s10_err.log || tail -5 gpurun_out/s10_err.log
done
python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --tile-bits 13 > gpurun_out/s10_c5_1.json 2>> gpurun_out/s10_err.log || tail -5 gpurun_out/s10_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s10_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], 'tile', d['config']['tile_bits'], round(d['ms_per_step'],1), 'ms')
PY
