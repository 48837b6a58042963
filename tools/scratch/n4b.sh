set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 4 --master-port 29581 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "rc=$?"
timeout 200 $TR --nproc-per-node 4 --master-port 29582 tools/timeline_dist.py --log2n 30 > gpurun_out/r02_timeline_n4_blocks.json 2> gpurun_out/r02_timeline_n4_blocks.err; echo "rc=$?"
timeout 200 $TR --nproc-per-node 2 --master-port 29583 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"ok": [a-z]*' gpurun_out/r02_bench_n4.json gpurun_out/r02_bench_n2.json
