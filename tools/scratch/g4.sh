set -x
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_g1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_ncu_g1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --worst > gpurun_out/r02_bench_g1.json 2> gpurun_out/r02_bench_g1.err
tail -c 3000 gpurun_out/r02_bench_g1.json
