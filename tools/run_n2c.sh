set -x
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/check_sharded_n2.log 2>&1; echo "check rc=$?"; grep -c OK gpurun_out/check_sharded_n2.log; grep -i "fail\|error" gpurun_out/check_sharded_n2.log | head
for c in 1 0; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$c bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --peer-copy $c > gpurun_out/bench_n2_p$c.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n2.err
done
