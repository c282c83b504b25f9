set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/ab8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab8_pytest.log
ms() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT $1', d['config']['workload'][:8], d['config']['qubits'], 'ms', round(d['ms_per_step'],2), 'passes', d['config']['passes'], 'parity', d.get('parity',{}).get('ok'))"; }
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/ab8_err.log | tee -a gpurun_out/ab8_lines.json | ms new
C="--workload c5 --qubits 28 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $C 2>>gpurun_out/ab8_err.log | tee -a gpurun_out/ab8_lines.json | ms new
PSUM_LIB_PATH=$PWD/.ab/lib_head.so python bench.py $C 2>>gpurun_out/ab8_err.log | tee -a gpurun_out/ab8_lines.json | ms head305
