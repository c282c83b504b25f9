set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_reference_run.py tests/test_gpu_shim.py -m gpu -x -q > gpurun_out/ab4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ab4_pytest.log
ms() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT $1', d['config']['workload'][:8], d['config']['qubits'], 'ms', round(d['ms_per_step'],2), 'passes', d['config']['passes'], 'sets', d['config']['partner_load_sets'], 'parity', d.get('parity',{}).get('ok'))"; }
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $B 2>>gpurun_out/ab4_err.log | tee -a gpurun_out/ab4_lines.json | ms default_rb3_tries64
python bench.py $B --plan-tries 8 2>>gpurun_out/ab4_err.log | tee -a gpurun_out/ab4_lines.json | ms tries8
python bench.py $B --reg-bits 2 2>>gpurun_out/ab4_err.log | tee -a gpurun_out/ab4_lines.json | ms rb2
python bench.py $B --reg-bits 2 --cta-mode 1 2>>gpurun_out/ab4_err.log | tee -a gpurun_out/ab4_lines.json | ms rb2_cta1
C="--workload c5 --qubits 28 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-parity"
python bench.py $C 2>>gpurun_out/ab4_err.log | tee -a gpurun_out/ab4_lines.json | ms c5_default
python bench.py $C --reg-bits 2 2>>gpurun_out/ab4_err.log | tee -a gpurun_out/ab4_lines.json | ms c5_rb2
python bench.py $C --min-cover 5 2>>gpurun_out/ab4_err.log | tee -a gpurun_out/ab4_lines.json | ms c5_mincover5
