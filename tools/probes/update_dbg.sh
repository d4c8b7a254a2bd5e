cd "$(dirname "$0")/../.."
python - <<'PY'
import sys
sys.path.insert(0, ".")
from tools import corpus
plain = corpus.text_like(64 << 20, seed=9)
open("/tmp/t64.lz4", "wb").write(corpus.build_frame(plain, 4, True, True))
open("/tmp/t64l.lz4", "wb").write(corpus.build_frame(plain, 4, True, True, independent=False))
open("/tmp/t64_4m.lz4", "wb").write(corpus.build_frame(plain, 7, False, True))
open("/tmp/t64.bin", "wb").write(plain)
PY
for f in /tmp/t64.lz4 /tmp/t64_4m.lz4 /tmp/t64l.lz4 tests/golden/z9m.lz4; do
echo "== $f --update --keep (one decompressor, 4 passes)"
LZ4ADA_UPDATE_DEBUG=1 ./tools/unlz4ada_b200 --update --keep -v --file $f --repeat 4 2>&1 >/tmp/o.bin | tail -4 | cut -c1-220
done
cmp /tmp/o.bin <(python -c "import sys; sys.stdout.buffer.write(bytes(9437166))") && echo "z9m ok"
./tools/unlz4ada_b200 --update --keep --file /tmp/t64.lz4 --repeat 2 | cmp - /tmp/t64.bin && echo "t64 keep ok"
