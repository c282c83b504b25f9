set -x
cd .old_wt
A="--workload c5 --qubits 28 --low-bits 3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
python bench.py $A > ../gpurun_out/ab_old_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_pass -c 90 --csv --log-file ../gpurun_out/ab_old_launches.csv python bench.py $A > ../gpurun_out/ab_old_ncu.log 2>&1
echo rc=$?
python bench.py --workload c5 --qubits 28 --low-bits 3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('OLD c5 28q ms', d['ms_per_step'], d['config']['passes'])"
python bench.py --qubits 28 --low-bits 3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('OLD c4 28q ms', d['ms_per_step'], d['config']['passes'])"
