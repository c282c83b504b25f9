set -x
for c in 0 8 16 32; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((c%10)) bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-parity --comm-sms $c > gpurun_out/bench_n2_c$c.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2_c$c.json').read().strip().splitlines()[-1])
n=d['nvlink']; print('comm_sms $c step', round(d['ms_per_step'],1), 'noex', round(n['step_ms_without_exchange'],1), 'ex_ms', round(n['exchange_ms_on_comm_stream'],1), 'gbs', round(n['achieved_gbs_per_direction']), 'overlap', round(n['overlap_frac'],2))
"
done
