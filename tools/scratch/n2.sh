set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29551 tests/dist_gpu_worker.py > gpurun_out/r02_dist_worker_n2.log 2>&1; echo "worker rc=$?"
grep "DIST_GPU" gpurun_out/r02_dist_worker_n2.log; tail -3 gpurun_out/r02_dist_worker_n2.log
timeout 300 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2_p2p.json 2> gpurun_out/r02_bench_n2_p2p.err; echo "bench p2p rc=$?"
URH_B200_P2P=0 timeout 300 $TR --nproc-per-node 2 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/r02_bench_n2_nccl.err; echo "bench nccl rc=$?"
grep -o '"exchange": "[^"]*"\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*' gpurun_out/r02_bench_n2_p2p.json gpurun_out/r02_bench_n2_nccl.json
tail -3 gpurun_out/r02_bench_n2_p2p.err
