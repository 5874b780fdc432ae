python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gputests_3.log
python tools/time_extract.py foa > gpurun_out/r2_time_foa_v4.log 2>&1
python tools/time_extract.py mic > gpurun_out/r2_time_mic_v4.log 2>&1
python tools/profile_extract.py --mode foa --clips 148 --iters 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:extract_kernel -s 2 -c 1 -f -o gpurun_out/r2_foa_v14 python tools/profile_extract.py --mode foa --clips 148 --iters 2 > gpurun_out/ncu_foa_v14.log 2>&1
python tools/profile_extract.py --mode mic --clips 148 --iters 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:extract_kernel -s 2 -c 1 -f -o gpurun_out/r2_mic_fused_v5 python tools/profile_extract.py --mode mic --clips 148 --iters 2 > gpurun_out/ncu_mic_v5.log 2>&1
cat gpurun_out/r2_gputests_3.log gpurun_out/r2_time_foa_v4.log gpurun_out/r2_time_mic_v4.log; tail -3 gpurun_out/ncu_foa_v14.log gpurun_out/ncu_mic_v5.log
