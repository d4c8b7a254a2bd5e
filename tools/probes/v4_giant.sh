mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "k1 or K1 or guard or batch or synthetic or mutation or corrupt" > gpurun_out/t_v4.log 2>&1; echo rc=$? >> gpurun_out/t_v4.log
tail -4 gpurun_out/t_v4.log
timeout 900 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-e2e --corpus-rank 3 > gpurun_out/bench_r3.json 2> gpurun_out/bench_r3.err; echo bench rc=$?
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r3.json').read().strip().splitlines()[-1])
print('headline', j['value'], j['ms_per_step'])
for c in j.get('configs',[]): print(c['name'][:60], round(c['value'],1), 'GB/s', round(c['ms_per_step'],2),'ms', c['kernel_ms'])
PY
