set -x
timeout 900 python -m pytest tests/test_gpu_circuits.py -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s13_pytest.log
timeout 600 python tools/time_stateprep.py 30 > gpurun_out/stateprep_30q.json 2> gpurun_out/s13_err.log; tail -3 gpurun_out/s13_err.log; cat gpurun_out/stateprep_30q.json
