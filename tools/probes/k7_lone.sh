cd "$(dirname "$0")/../.."
python - <<'PY'
import sys
sys.path.insert(0, ".")
from tools import corpus
plain = corpus.text_like(64 << 20, seed=9)
open("/tmp/t64l.lz4", "wb").write(corpus.build_frame(plain, 4, True, True, independent=False))
open("/tmp/t64_4m.lz4", "wb").write(corpus.build_frame(plain, 7, False, True))
PY
for f in /tmp/t64l.lz4 /tmp/t64_4m.lz4; do
  echo "== $f"; ./tools/unlz4ada_b200 -v --file $f --repeat 3 2>&1 >/dev/null | tail -2 | cut -c1-200
  echo "== $f --update"; ./tools/unlz4ada_b200 --update -v --file $f --repeat 3 2>&1 >/dev/null | tail -1 | cut -c1-200
done
