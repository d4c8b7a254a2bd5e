cd "$(dirname "$0")/../.."
python - <<'PY'
import sys
sys.path.insert(0, ".")
from tools import corpus
plain = corpus.text_like(64 << 20, seed=9)
open("/tmp/t64.lz4", "wb").write(corpus.build_frame(plain, 4, True, True))
PY
LZ4ADA_UPDATE_DEBUG=1 ./tools/unlz4ada_b200 --update --keep -v --file /tmp/t64.lz4 --repeat 4 2>&1 >/dev/null | tail -3 | cut -c1-300
