set -x
timeout 1200 python -m pytest tests/test_gpu_eigensolver.py tests/test_gpu_shim.py tests/test_reference_run.py tests/test_matrix_operator.py -m gpu -x -q > gpurun_out/n3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/n3_pytest.log
A="--qubits 26 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $A > gpurun_out/n3_plain26.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pass -c 4 -o gpurun_out/prof_r2c_q26 python bench.py $A > gpurun_out/n3_ncu26.log 2>&1
echo rc=$?
