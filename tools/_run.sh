python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gputests_6.log
python tools/time_extract.py mic > gpurun_out/r2_time_mic_v6.log 2>&1
python tools/time_extract.py foa > gpurun_out/r2_time_foa_v7.log 2>&1
cat gpurun_out/r2_gputests_6.log gpurun_out/r2_time_mic_v6.log gpurun_out/r2_time_foa_v7.log
