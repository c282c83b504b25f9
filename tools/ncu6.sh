set -x
A="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $A > gpurun_out/n6_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_pass -c 14 --csv --log-file gpurun_out/n6_launches.csv python bench.py $A > gpurun_out/n6_ncu.log 2>&1
echo rc=$?
