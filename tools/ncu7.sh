set -x
A="--reg-bits 4 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $A > gpurun_out/n7_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_pass -c 14 --csv --log-file gpurun_out/n7_launches.csv python bench.py $A > gpurun_out/n7_ncu.log 2>&1
echo rc=$?
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/s12_c4_1.json 2> gpurun_out/s12_err.log || tail -5 gpurun_out/s12_err.log
python bench.py --workload c5 --qubits 30 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/s12_c5_1.json 2>> gpurun_out/s12_err.log || tail -5 gpurun_out/s12_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s12_c*.json')):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config']['workload'][:2], 'passes', d['config']['passes'], round(d['ms_per_step'],1), 'ms')
PY
