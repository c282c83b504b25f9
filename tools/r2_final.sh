set -x
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/f_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"
A="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -c 200 --csv --log-file gpurun_out/f_launches.csv python bench.py $A > gpurun_out/f_ncu.log 2>&1; echo "ncu launches rc=$?"
B="--qubits 26 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $B > gpurun_out/f_plain26.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 12 -f -o gpurun_out/prof_r2_final_q26 python bench.py $B > gpurun_out/f_ncu26.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/f_bench.json | cut -c1-700
