set -x
SECONDS=0
python bench.py > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
echo "bench wall $SECONDS s" > gpurun_out/r2_bench_n1_b.wall
SECONDS=0
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_ref_b.json 2> gpurun_out/r2_bench_ref_b.err
echo "ref wall $SECONDS s" >> gpurun_out/r2_bench_n1_b.wall
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-config5 > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/r2_bench_n1_b.wall; tail -c 600 gpurun_out/r2_bench_n1_b.err
