set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 4 --master-port 29571 tools/timeline_dist.py --log2n 30 > gpurun_out/r02_timeline_n4_nccl.json 2> gpurun_out/r02_timeline_n4_nccl.err; echo "rc=$?"
URH_B200_P2P=1 timeout 200 $TR --nproc-per-node 4 --master-port 29572 tools/timeline_dist.py --log2n 30 > gpurun_out/r02_timeline_n4_mailboxes.json 2> gpurun_out/r02_timeline_n4_mailboxes.err; echo "rc=$?"
tail -c 400 gpurun_out/r02_timeline_n4_nccl.json; tail -3 gpurun_out/r02_timeline_n4_nccl.err
