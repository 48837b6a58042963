set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 20 --warmup 3 --no-parity --no-e2e --no-cpu"
timeout 200 $TR --nproc-per-node 2 --master-port 29561 $B > gpurun_out/x_all.json 2>/dev/null
URH_B200_P2P_NO_REDUCE=1 timeout 200 $TR --nproc-per-node 2 --master-port 29562 $B > gpurun_out/x_noreduce.json 2>/dev/null
URH_B200_P2P_NO_GATHER=1 timeout 200 $TR --nproc-per-node 2 --master-port 29563 $B > gpurun_out/x_nogather.json 2>/dev/null
URH_B200_P2P=0 timeout 200 $TR --nproc-per-node 2 --master-port 29564 $B > gpurun_out/x_nccl.json 2>/dev/null
grep -o '"ms_per_step": [0-9.]*' gpurun_out/x_all.json gpurun_out/x_noreduce.json gpurun_out/x_nogather.json gpurun_out/x_nccl.json
