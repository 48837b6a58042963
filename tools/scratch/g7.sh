set -x
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 600 python tools/bench_paths.py > gpurun_out/r02_paths_fifo.jsonl 2> gpurun_out/r02_paths_fifo.err
URH_B200_FSK_NO_FIFO=1 timeout 600 python tools/bench_paths.py > gpurun_out/r02_paths_nofifo.jsonl 2> gpurun_out/r02_paths_nofifo.err
grep "int16\|int8\|afp_demod FSK" gpurun_out/r02_paths_fifo.jsonl gpurun_out/r02_paths_nofifo.jsonl
tail -3 gpurun_out/r02_paths_fifo.err
