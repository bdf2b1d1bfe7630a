mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/s5_tests11.log 2>&1
tail -3 gpurun_out/s5_tests11.log
(timeout 300 python bench.py --config c3q25 --steps 2 --warmup 2 --no-cpu-baseline) > gpurun_out/s5_bench_q25b.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/s5_bench_q25b.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['phase_ms'])
"
