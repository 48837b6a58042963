set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8
timeout 240 $TR --nproc-per-node 8 --master-port 29541 tests/dist_gpu_worker.py > gpurun_out/r02_dist_worker_n8.log 2>&1; echo "worker rc=$?"
tail -3 gpurun_out/r02_dist_worker_n8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8 rc=$?"
timeout 240 $TR --nproc-per-node 8 --master-port 29544 tools/bench_config4.py --log2n 30 --steps 3 --given-only > gpurun_out/r02_config4_n8.json 2> gpurun_out/r02_config4_n8.err; echo "config4 rc=$?"
tail -c 1500 gpurun_out/r02_bench_n8.json; tail -c 1500 gpurun_out/r02_config4_n8.json; tail -5 gpurun_out/r02_bench_n8.err gpurun_out/r02_config4_n8.err
