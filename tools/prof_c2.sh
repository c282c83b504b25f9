set -x
python tools/prof_c2.py > gpurun_out/n8_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pass -c 12 -o gpurun_out/prof_r2e_c2_q24 python tools/prof_c2.py > gpurun_out/n8_ncu.log 2>&1
echo rc=$?
cd .old_wt && python ../tools/prof_c2.py > ../gpurun_out/n8_old_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_pass -c 24 --csv --log-file ../gpurun_out/n8_old_launches.csv python ../tools/prof_c2.py > ../gpurun_out/n8_old_ncu.log 2>&1
echo rc=$?
