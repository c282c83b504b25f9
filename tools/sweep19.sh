set -x
A="--steps 3 --warmup 3 --no-cpu-baseline"
python bench.py $A --host-chunk-bits 5 > gpurun_out/s19_e2e_1.json 2> gpurun_out/s19_err.log
python bench.py $A --host-chunk-bits 5 --copy-threads 15 > gpurun_out/s19_e2e_2.json 2>> gpurun_out/s19_err.log
python bench.py $A --copy-threads 8 > gpurun_out/s19_e2e_3.json 2>> gpurun_out/s19_err.log
python bench.py $A --host-chunk-bits 6 > gpurun_out/s19_e2e_4.json 2>> gpurun_out/s19_err.log
tail -5 gpurun_out/s19_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s19_*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']; print(f.split('/')[-1], round(d['ms_per_step'],1), 'ms', 'e2e pageable', round(e['ms_per_step'],1), 'pinned', round(e['pinned']['ms_per_step'],1), 'device expect', round(e['device_expect_ms'],1), 'parity', d['parity']['ok'], d['parity']['rel_err_energy_e2e_vs_device'])
PY
