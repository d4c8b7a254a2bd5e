#!/bin/bash
# Throughput of the drop-in Update API and of the batch entry through the C tool (no Python in the timed path):
# a 64 MiB text frame (64 KiB independent blocks, block + content checksum), the same linked, and z9m.
set -e
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
for f in /tmp/t64.lz4 /tmp/t64l.lz4 /tmp/t64_4m.lz4 tests/golden/z9m.lz4; do
  echo "== $f --update --keep (ONE decompressor for 4 passes: passes 2.. are the steady state of a long stream)"
  ./tools/unlz4ada_b200 --update --keep -v --file $f --repeat 4 2>&1 >/dev/null | tail -3
  for mode in "--update" ""; do
    echo "== $f $mode (3 passes in one process; the first pays for the CUDA context)"
    ./tools/unlz4ada_b200 $mode -v --file $f --repeat 3 2>&1 >/tmp/out.bin | tail -3
  done
done
cmp /tmp/out.bin <(python -c "import sys; sys.stdout.buffer.write(bytes(9437166))") && echo "z9m output ok"
./tools/unlz4ada_b200 < /tmp/t64.lz4 | cmp - /tmp/t64.bin && echo "t64 batch output ok"
./tools/unlz4ada_b200 --update < /tmp/t64l.lz4 | cmp - /tmp/t64.bin && echo "t64 linked update output ok"
