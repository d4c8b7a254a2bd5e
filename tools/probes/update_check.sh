cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "update or Update or read_ahead or stream or cli or vector or lifetime" > gpurun_out/t_upd.log 2>&1; echo rc=$? >> gpurun_out/t_upd.log
tail -3 gpurun_out/t_upd.log
python - <<'PY'
import sys
sys.path.insert(0, ".")
from tools import corpus
plain = corpus.text_like(64 << 20, seed=9)
open("/tmp/t64.lz4", "wb").write(corpus.build_frame(plain, 4, True, True))
PY
for f in /tmp/t64.lz4 tests/golden/z9m.lz4; do
echo "== $f --update, a new decompressor per pass (5 passes)"; ./tools/unlz4ada_b200 --update -v --file $f --repeat 5 2>&1 >/dev/null | tail -3 | cut -c1-200
done
