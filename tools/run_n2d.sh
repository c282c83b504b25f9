set -x
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/check_sharded_n2.log 2>&1; echo "check rc=$?"; grep -c OK gpurun_out/check_sharded_n2.log; grep -i "fail\|error" gpurun_out/check_sharded_n2.log | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('n2', round(d['ms_per_step'],1), 'ms value', round(d['value']), 'e2e', e.get('ms_per_step'), (e.get('pinned') or {}).get('ms_per_step'), 'parity', d['parity']['ok'], 'nvlink', d['nvlink']['frac'], d['nvlink']['overlap_frac'])
PY
