set -x
A="--workload c5 --qubits 26 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $A > gpurun_out/n4_plain26.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pass -s 3 -c 3 -o gpurun_out/prof_r2d_c5_q26 python bench.py $A > gpurun_out/n4_ncu26.log 2>&1
echo rc=$?
