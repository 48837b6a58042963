set -x
timeout 600 python -m pytest tests/test_gpu_onecall.py tests/test_gpu_demod_digitize.py -q -x 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_g2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_ncu_g2.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_g2.json 2> gpurun_out/r02_bench_g2.err
grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"ok": [a-z]*' gpurun_out/r02_bench_g2.json
