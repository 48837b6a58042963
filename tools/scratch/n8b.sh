set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8 rc=$?"
URH_B200_P2P=1 timeout 200 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 3 --no-parity --no-e2e --no-cpu > gpurun_out/r02_bench_n8_p2p.json 2> gpurun_out/r02_bench_n8_p2p.err; echo "bench8 p2p rc=$?"
timeout 200 $TR --nproc-per-node 4 --master-port 29545 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench4 rc=$?"
grep -o '"exchange": "[^"]*"\|"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"ok": [a-z]*' gpurun_out/r02_bench_n8.json gpurun_out/r02_bench_n8_p2p.json gpurun_out/r02_bench_n4.json
tail -2 gpurun_out/r02_bench_n8_p2p.err
