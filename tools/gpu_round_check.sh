#!/usr/bin/env bash
# The measurement set behind profiles/r2b_*, as run on the GPU box:  gpurun --timeout 3000 -- 'bash tools/gpu_round_check.sh'
# Everything lands in gpurun_out/ (scratch); the summaries that are judged are copied into profiles/ by hand.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests -m gpu -q --timeout=300 ) > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2b_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r2b_bench_full_n1.json 2> gpurun_out/r2b_bench_full_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_bench_reference_arm.err; echo "reference arm rc=$?"
# launch list of the bench command (2 windows + the DiT stage: ncu serialises every launch) and full captures of the three tensor / token-loop kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --windows 2 --steps 1 --warmup 1 --no-cpu-baseline --oracle-check none > gpurun_out/r2b_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_launches.csv "bench.py --windows 2 (decode + DiT), every launch" > gpurun_out/r2b_launches.md 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel_ll -s 2 -c 1 -f -o gpurun_out/r2b_mega_ll \
    python bench.py --windows 4 --steps 1 --warmup 1 --no-cpu-baseline --oracle-check none --dit 0 > gpurun_out/r2b_ncu_mega.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_tc_kernel|gemm_tf32x3_kernel" -s 80 -c 12 -f -o gpurun_out/r2b_tensor_kernels \
    python tools/encode_one.py 2 16 > gpurun_out/r2b_ncu_tensor.log 2>&1
# timelines and micro-benchmarks
timeout 300 python tools/mega3_trace.py > gpurun_out/r2b_mega_trace.txt 2>&1
timeout 300 python tools/stage_times.py > gpurun_out/r2b_stage_times.txt 2>&1
timeout 300 python tools/dit_bench.py 1 > gpurun_out/r2b_dit_bench.txt 2>&1
timeout 300 python tools/attn_tc_bench.py > gpurun_out/r2b_attn_tc_bench.txt 2>&1
# every song seed the 8-GPU scaling run will use (rank r decodes song r): decode only, self-consistency + token stream hash
for seed in 1 2 3 4 5 6 7; do
  timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-check none --dit 0 --windows 80 --song-seed $seed 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({'song_seed': d.get('song_seed'), 'value': d.get('value'), 'e2e': d.get('e2e',{}).get('value'), 'self_consistency': d.get('self_consistency'), 'token_stream_sha1': d.get('token_stream_sha1')}))
" >> gpurun_out/r2b_seed_sweep.jsonl
done
tail -n 3 gpurun_out/r2b_*.txt; cat gpurun_out/r2b_seed_sweep.jsonl; head -c 1500 gpurun_out/r2b_bench_full_n1.json
