mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_enc1.csv python tools/encode_one.py 2 > gpurun_out/enc1.log 2>&1; wc -l gpurun_out/launches_enc1.csv
