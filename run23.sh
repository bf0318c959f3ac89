set -x
timeout 300 python tools/mega_trace.py 2>&1 | tail -14
timeout 600 python tools/attn_bench.py 1 8 32 64 2>&1 | tail -6
