python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gputests_9.log
python tools/time_extract.py mic > gpurun_out/r2_time_mic_v9.log 2>&1
cat gpurun_out/r2_gputests_9.log gpurun_out/r2_time_mic_v9.log
