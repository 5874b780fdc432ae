python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8_d.json 2> gpurun_out/r2_bench_n8_d.err
tail -c 200 gpurun_out/r2_bench_n8_d.err; wc -c gpurun_out/r2_bench_n8_d.json
