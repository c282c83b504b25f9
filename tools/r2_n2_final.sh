set -x
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/n2f_bench.json 2> gpurun_out/n2f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/n2f_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/n2f_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('RESULT n2', round(d['ms_per_step'],1), 'ms value', round(d['value']), 'passes', d['config']['passes'], 'parity', d['parity']['ok'], d['parity'].get('max_rel_err_y'), 'lanczos', d['parity'].get('sharded_lanczos',{}).get('max_abs_err_vs_single_gpu'), 'nvlink', d['nvlink']['frac'], d['nvlink']['overlap_frac'])
PY
