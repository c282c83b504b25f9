set -x
python bench.py --qubits 26 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/n1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pass -c 4 -o gpurun_out/prof_r2a_q26 \
    python bench.py --qubits 26 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/n1_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/n1_ncu.log
