cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for cfg in "8 29" "16 28" "32 27" "12 28"; do
  set -- $cfg
  LZ4ADA_E2E_CHUNKS=$1 LZ4ADA_E2E_CHUNK_SHIFT=$2 timeout 300 python bench.py --steps 2 --warmup 3 --skip-configs --skip-cpu-baseline --e2e-steps 4 > gpurun_out/e2e_$1.json 2>/dev/null
  python -c "
import json
j=json.loads(open('gpurun_out/e2e_$1.json').read().strip().splitlines()[-1]); e=j['e2e']
print('chunks $1', round(e['value'],2), 'GB/s', round(e['ms_per_step'],1), 'ms; ceiling', round(e['copy_ceiling_ms'],1), e['e2e_k1_kernel'])"
done
