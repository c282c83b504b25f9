set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for args in "" "--low-bits 3" "--tile-bits 13" "--tile-bits 13 --low-bits 3" "--low-bits 2"; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $args > gpurun_out/s1_$(echo $args | tr -d ' -').json 2> gpurun_out/s1_err.log || tail -5 gpurun_out/s1_err.log
done
for args in "" "--low-bits 3" "--tile-bits 13 --low-bits 3"; do
  python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e $args > gpurun_out/s1_c5_$(echo $args | tr -d ' -').json 2>> gpurun_out/s1_err.log || tail -5 gpurun_out/s1_err.log
done
cat gpurun_out/s1_*.json | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['config']['workload'], d['config']['passes'], d['config']['tile_bits'], round(d['ms_per_step'],1))
"
