mkdir -p gpurun_out
P="python tools/k1_probe.py --size-mib 512 --frame-mib 2 --block 256k --linked --kinds text --tunings 0 --reps 1"
echo "== k7 no block checksum"; LZ4B200_K7_DEBUG=1 timeout 120 $P --no-block-checksum 2>&1 | tail -3 | cut -c1-400
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "chain_kernel_variants and k7" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --e2e-steps 2 > gpurun_out/bench_k7.json 2> gpurun_out/bench_k7.err; echo bench rc=$?
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_k7.json').read().strip().splitlines()[-1])
print('headline', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'])
for c in j.get('configs',[]): print(c['name'][:60], round(c['value'],1), 'GB/s', round(c['ms_per_step'],2),'ms', c['kernel_ms'])
PY
