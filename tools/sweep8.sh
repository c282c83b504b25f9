set -x
i=0
for args in "" "--share-weight 0" "--share-weight 30" "--share-weight 300"; do
  i=$((i+1))
  python bench.py --qubits 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity $args > gpurun_out/s8_c4_$i.json 2> gpurun_out/s8_err.log || tail -5 gpurun_out/s8_err.log
  python bench.py --workload c5 --qubits 28 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity $args > gpurun_out/s8_c5_$i.json 2>> gpurun_out/s8_err.log || tail -5 gpurun_out/s8_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s8_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], 'sets', d['config']['partner_load_sets'], round(d['ms_per_step'],1), 'ms')
PY
