mkdir -p gpurun_out
for tc in 1 0; do
timeout 900 python bench.py --windows 48 --steps 1 --warmup 3 --no-cpu-baseline --tc $tc > gpurun_out/bench_tc$tc.json 2> gpurun_out/bench_tc$tc.err; tail -2 gpurun_out/bench_tc$tc.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc$tc.json')); print('tc',$tc,'value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], d['token_stream_sha1'])"
done
