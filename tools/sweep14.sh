set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s14_pytest.log
timeout 600 python tools/time_batch.py > gpurun_out/batch_24q.json 2> gpurun_out/s14_err.log; tail -3 gpurun_out/s14_err.log; cat gpurun_out/batch_24q.json
