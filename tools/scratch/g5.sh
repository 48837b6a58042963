set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fsk_fifo|k_hist_interior_dev|k_dense_f32|k_finish_rows" -s 10 -c 9 -f -o gpurun_out/r02_prof_final python bench.py --log2n 28 --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_ncu_final.log 2>&1
ls -la gpurun_out/r02_prof_final.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_ncu_g1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --worst > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_refarm.json 2> gpurun_out/r02_bench_refarm.err
tail -c 600 gpurun_out/r02_bench_refarm.json
timeout 600 python tools/bench_paths.py > gpurun_out/r02_secondary.jsonl 2> gpurun_out/r02_secondary.err
timeout 600 python tools/bench_configs.py > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
tail -2 gpurun_out/r02_configs.err
