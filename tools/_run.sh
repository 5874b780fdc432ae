python bench.py > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err
tail -c 200 gpurun_out/r2_bench_n1_e.err
python tools/profile_extract.py --mode mic --clips 148 --iters 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:extract_kernel -s 2 -c 1 -f -o gpurun_out/r2_mic_fused_v8 python tools/profile_extract.py --mode mic --clips 148 --iters 2 > gpurun_out/ncu_mic_v8.log 2>&1
tail -2 gpurun_out/ncu_mic_v8.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:extract_kernel|stats_|finalize|mask_kernel|augment|clip_max' -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r2_ncu_launches_bench.csv
