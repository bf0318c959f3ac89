set -x
timeout 120 ./build_tools/l2_hotspot_bench
timeout 300 python tools/mega_trace.py 2>&1 | tail -13
