python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8_c.json 2> gpurun_out/r2_bench_n8_c.err
tail -c 300 gpurun_out/r2_bench_n8_c.err; wc -c gpurun_out/r2_bench_n8_c.json
