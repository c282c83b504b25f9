set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3_pytest.log
B="--qubits 26 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $B > gpurun_out/s3_plain26.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 6 -f -o gpurun_out/prof_s3_q26 python bench.py $B > gpurun_out/s3_ncu26.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
