set -x
A="--low-bits 3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $A > gpurun_out/n2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:k_pass -c 14 --csv --log-file gpurun_out/n2_launches.csv python bench.py $A > gpurun_out/n2_ncu.log 2>&1
echo rc=$?
python bench.py --qubits 26 $A > gpurun_out/n2_plain26.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pass -c 3 -o gpurun_out/prof_r2b_q26 python bench.py --qubits 26 $A > gpurun_out/n2_ncu26.log 2>&1
echo rc=$?
