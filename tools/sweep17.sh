set -x
python tools/h2d_probe.py > gpurun_out/s17_h2d.json 2> gpurun_out/s17_err.log; cat gpurun_out/s17_h2d.json
A="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $A --low-bits 2 > gpurun_out/s17_c4_1.json 2>> gpurun_out/s17_err.log
python bench.py $A --low-bits 1 > gpurun_out/s17_c4_2.json 2>> gpurun_out/s17_err.log
python bench.py $A --low-bits 2 --tile-bits 13 > gpurun_out/s17_c4_3.json 2>> gpurun_out/s17_err.log
C="--workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $C --low-bits 2 > gpurun_out/s17_c5_1.json 2>> gpurun_out/s17_err.log
python bench.py $C --low-bits 2 --tile-bits 13 > gpurun_out/s17_c5_2.json 2>> gpurun_out/s17_err.log
python bench.py $C --low-bits 1 --tile-bits 13 > gpurun_out/s17_c5_3.json 2>> gpurun_out/s17_err.log
tail -5 gpurun_out/s17_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s17_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], 'tile', d['config']['tile_bits'], round(d['ms_per_step'],1), 'ms', 'dram_frac', round(d['roofline']['dram_frac'],3))
PY
