set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python tools/mega_trace.py 2>&1 | tail -13
python tools/stage_times.py 2>&1 | tail -5
