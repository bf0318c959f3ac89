set -x
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/stage_times.py 2>&1 | tail -5
MB200_SMALLM=1 python tools/stage_times.py 2>&1 | tail -3
