set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29551 tests/dist_gpu_worker.py > gpurun_out/r02_dist_worker_n2.log 2>&1; echo "worker rc=$?"
grep "DIST_GPU" gpurun_out/r02_dist_worker_n2.log
bash tools/scratch/g1.sh
