python tools/time_extract.py foa > gpurun_out/r2_time_foa_v6.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_gputests_5.log
cat gpurun_out/r2_time_foa_v6.log gpurun_out/r2_gputests_5.log
