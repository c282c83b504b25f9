set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/check_sharded_n8.log 2>&1; echo "check rc=$?"; grep -c OK gpurun_out/check_sharded_n8.log; grep -i "fail\|error" gpurun_out/check_sharded_n8.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload c5 --qubits 31 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_n8_c5_34q.json 2> gpurun_out/bench_n8_c5.err; echo "bench c5 rc=$?"; tail -3 gpurun_out/bench_n8_c5.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 rc=$?"; tail -3 gpurun_out/bench_n4.err
