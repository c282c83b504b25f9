set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/b1.json 2> gpurun_out/b1.err; echo rc=$?; tail -3 gpurun_out/b1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b1_ref.json 2>> gpurun_out/b1.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','parity','e2e','roofline','cpu_baseline'): print(k, json.dumps(d.get(k))[:900])
"
