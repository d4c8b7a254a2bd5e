"""CPU-only checks of the product library: the C-ABI loads and exports every symbol that
include/lz4b200.h declares, and the host layer (frame-header parser, block-table builder, host
XXHash32, exception texts) agrees with the oracle / the reference's golden vectors.  No kernel is
launched here; calls that would need the GPU must fail loudly instead of falling back."""
import ctypes
import json
import os
import re

import pytest

import bo_lz4_ada_b200 as lz
from oracle_binding import OracleError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
INLINE = json.load(open(os.path.join(GOLDEN, "inline_cases.json")))


def _read(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "lz4b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lz4(?:b200|ada)_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    handle = ctypes.CDLL(lz.SO_PATH)
    declared = _declared_functions()
    assert len(declared) >= 50
    missing = [n for n in declared if not hasattr(handle, n)]
    assert not missing, missing
    # the python binding covers the same set
    assert sorted(lz.exported_symbols()) == declared


def test_no_oracle_or_cpu_decoder_in_product_library():
    """The product must not link the oracle or any LZ4 CPU decoder."""
    import subprocess
    out = subprocess.run(["nm", "-D", lz.SO_PATH], capture_output=True, text=True).stdout
    assert "lzo_" not in out and "LZ4_decompress" not in out
    ldd = subprocess.run(["ldd", lz.SO_PATH], capture_output=True, text=True).stdout
    assert "liblz4" not in ldd and "oracle" not in ldd


def test_host_xxhash32_kat_and_oracle(oracle):
    case = INLINE["xxh32_kat"]
    h = lz.XXHash32.Init()
    for b in bytes.fromhex(case["input_hex"]):
        h.Update(bytes([b]))
    assert "%08x" % h.Final() == case["expect"]
    import numpy as np
    rng = np.random.default_rng(3)
    for n in [0, 1, 3, 15, 16, 17, 31, 32, 33, 100, 1000, 65536, 99999]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert lz.XXHash32.Hash(data) == oracle.xxh32(data)
        h = lz.XXHash32.Init()
        for i in range(0, n, 7):
            h.Update(data[i:i + 7])
        assert h.Final() == oracle.xxh32(data)
    # Init ignores its seed (lib/lz4ada.adb:925-930), Reset honours it
    assert lz.XXHash32.Init(5).Final() == lz.XXHash32.Init(0).Final()
    h = lz.XXHash32.Init()
    h.Reset(5)
    assert h.Final() != lz.XXHash32.Init(0).Final()


def test_to_hex():
    assert lz.To_Hex(0xA7, 8) == "a7" and lz.To_Hex(0x184D9904) == "184d9904"


HEADER_ONLY_ERRORS = ["t2e", "corruptedmagic", "z1ver", "corruptedreserved", "corruptedblocksz", "corruptedhdrchck"]


@pytest.mark.parametrize("stem", HEADER_ONLY_ERRORS)
def test_header_error_vectors_init_with_header(stem):
    """Errors raised by Init_With_Header itself (lz4test.adb:284): no block, no device needed."""
    data = _read(stem + ".err")[:10001]
    with pytest.raises(lz.LZ4AdaError) as ei:
        lz.Init_With_Header(data, "Single_Frame")
    assert ei.value.information == MAN["error"][stem]["eds"]


def test_size_word_limit_error_needs_no_device():
    """cntblkszoverflow: the size-word check (lib/lz4ada.adb:541-553) fires before any decode."""
    data = _read("cntblkszoverflow.err")[:10001]
    ctx, consumed = lz.Init_With_Header(data, "Single_Frame")
    with pytest.raises(lz.Data_Corruption) as ei:
        while consumed < len(data):
            c, out, of, ol = ctx.Update(data[consumed:])
            consumed += c
    assert ei.value.information == MAN["error"]["cntblkszoverflow"]["eds"]


def test_reservation_exceeded_and_messages():
    case = INLINE["reservation_exceeded"]
    with pytest.raises(lz.Too_Little_Memory) as ei:
        lz.Init_With_Header(bytes.fromhex(case["input_hex"]), "SZ_64_KiB")
    assert "requres reservation SZ_1_MIB" in str(ei.value) and "only SZ_64_KIB be used" in str(ei.value)
    with pytest.raises(lz.Assertion_Error):
        lz.Init_With_Header(b"\x04\x22\x4d", "Single_Frame")   # Pre => Input'Length >= 7


def test_min_buffer_size_and_eof_defaults():
    for res, bm in [("SZ_64_KiB", 65536), ("SZ_4_MiB", 4 << 20), ("For_All", 8 << 20)]:
        d = lz.Init(res)
        assert d.Min_Buffer_Size == bm + 65536 + 8      # lib/lz4ada.adb:54
        assert d.Is_End_Of_Frame() == "No"
    d = lz.Init_For_Block(14)
    assert d.Is_End_Of_Frame() == "No"


def test_header_bytes_one_at_a_time_match_oracle(oracle):
    """Header-stage stepping: consumed counts per call equal the oracle's."""
    for stem in ["z100", "t2", "skippable", "z100legacy", "emptycraft"]:
        data = _read(stem + ".lz4")
        a, b = lz.Init(), oracle.init()
        # feed only bytes that cannot complete a block: stop at the first call that would decode
        for i in range(min(len(data), 8)):
            try:
                ca = a.Update(data[i:i + 1])[0]
            except lz.Device_Error:
                break
            cb = b.update(data[i:i + 1])[0]
            assert ca == cb == 1
            assert a.Is_End_Of_Frame() == b.is_end_of_frame()


def _plan(data_list, reservation="For_All"):
    src = b"".join(data_list)
    offs, pos = [], 0
    for d in data_list:
        offs.append((pos, len(d)))
        pos += len(d)
    return lz.Batch(None, src or b"\0", offs, reservation), src


def _expected_blocks(frame):
    """Independent walk of a single modern frame's size words (test-side)."""
    import struct
    flg = frame[4]
    pos = 7 + (8 if flg & 8 else 0) + (4 if flg & 1 else 0)
    bchk = 4 if flg & 0x10 else 0
    blocks = []
    while True:
        word = struct.unpack_from("<I", frame, pos)[0]
        pos += 4
        if word == 0:
            break
        n = word & 0x7FFFFFF
        blocks.append((pos, n, bool(word & 0x80000000)))
        pos += n + bchk
    return blocks


@pytest.mark.parametrize("stem", ["t300k", "t301k", "b3444k", "z9m", "z2841", "t1111k", "z1", "empty"])
def test_block_table_builder(stem):
    """The planner's block table = the frame's size words (offset, size, stored flag, checksum flag)."""
    data = _read(stem + ".lz4")
    b, _ = _plan([data])
    exp = _expected_blocks(data)
    assert b.block_count == len(exp)
    flg = data[4]
    for i, (off, n, stored) in enumerate(exp):
        d = b.block_desc(i)
        assert (d.src_off, d.src_len) == (off, n)
        assert bool(d.flags & 1) == stored
        assert bool(d.flags & 2) == bool(flg & 0x10)
        linked = not (flg & 0x20) and len(exp) > 1
        # big compressed blocks of independent frames (>= 64 KiB compressed) are placed as chains of one block
        # (LZ4B200_BLK_SOLO) for the chain kernel while a batch holds few of them; LZ4B200_SOLO=0 turns that off
        solo = os.environ.get("LZ4B200_SOLO", "") != "0" and not linked and not stored and n >= 65536
        assert bool(d.flags & 32) == solo             # LZ4B200_BLK_SOLO
        assert bool(d.flags & 8) == (linked or solo)  # LZ4B200_BLK_CHAINED
        assert bool(d.flags & 16) == (i == 0 or solo) # LZ4B200_BLK_FIRST_OF_FRAME
    ho = b.host_outcome(0)
    assert ho["exception"] == "OK" and ho["end_of_frame"] == "Yes" and ho["n_blocks"] == len(exp)


def test_planner_host_errors_match_oracle_streaming(oracle):
    """Host-detectable errors (headers, size words) and the EOF tri-state of the planner equal what
    the oracle's Init(For_All)+Update loop reports for the same stream."""
    cases = ["corruptedmagic.err", "z1ver.err", "corruptedreserved.err", "corruptedblocksz.err",
             "corruptedhdrchck.err", "z100legacy.lz4", "concatlegacy.lz4", "skippable.lz4", "skipz100.lz4",
             "z101legacyplus.lz4", "concat390.lz4", "t2e.err"]
    streams = [_read(c) for c in cases]
    b, _ = _plan(streams)
    for k, c in enumerate(cases):
        exc, out, eof, msg = oracle.decode_stream(streams[k], chunk=0, out_cap=1 << 20)
        ho = b.host_outcome(k)
        if exc in ("NOT_SUPPORTED", "TOO_LITTLE_MEMORY") or c == "corruptedhdrchck.err":
            assert ho["exception"] == exc and ho["message"] == msg, c
        else:
            assert ho["exception"] == "OK", (c, ho)
            assert ho["end_of_frame"] == eof, (c, ho, eof)


def test_planner_init_with_header_semantics(oracle):
    """Reservation Single_Frame / Use_First plans every stream as Init_With_Header(stream) + Update
    (lib/lz4ada.adb:79-125): the host-detectable .err vectors give their .eds line (first 10 001 bytes, like
    Test_Error_Case), concatenated frames trip the Single_Frame policing, a stream shorter than 7 bytes violates
    the precondition, and a block table comes out for the good ones -- all without a device."""
    host_cases = ["corruptedmagic", "z1ver", "corruptedreserved", "corruptedblocksz", "corruptedhdrchck", "t2e",
                  "cntblkszoverflow"]
    streams = [_read(c + ".err")[:10001] for c in host_cases]
    extra = [_read("concatlegacy.lz4"), _read("concat390.lz4"), _read("t300k.lz4"), b"\x04\x22\x4d"]
    b = lz.Batch(None, b"".join(streams + extra), _offsets(streams + extra), Reservation="Single_Frame")
    for k, c in enumerate(host_cases):
        ho = b.host_outcome(k)
        assert ho["message"] == MAN["error"][c]["eds"], c
    ho = b.host_outcome(len(host_cases))
    assert ho["exception"] == "DATA_CORRUPTION" and "looks like the beginning of another frame" in ho["message"]
    ho = b.host_outcome(len(host_cases) + 1)
    # modern + modern: the first frame must be decoded before the policing fires (content checksum first), so the
    # host stage records the error; it is reported unless a device-side error comes earlier in stream order
    assert ho["exception"] == "DATA_CORRUPTION" and "data was provided after End of Frame" in ho["message"]
    ho = b.host_outcome(len(host_cases) + 2)
    assert ho["exception"] == "OK" and ho["n_blocks"] == 5 and ho["end_of_frame"] == "Yes"
    ho = b.host_outcome(len(host_cases) + 3)
    assert ho["exception"] == "ASSERTION_ERROR"
    b.close()
    # Use_First: no policing, concatenated frames are walked to the end
    b = lz.Batch(None, b"".join(extra[:2]), _offsets(extra[:2]), Reservation="Use_First")
    for k in range(2):
        ho = b.host_outcome(k)
        assert ho["exception"] == "OK" and ho["n_frames"] == 2, ho
    b.close()


def _offsets(streams):
    offs, pos = [], 0
    for s in streams:
        offs.append((pos, len(s)))
        pos += len(s)
    return offs


def test_planner_truncated_streams(oracle):
    """Streams cut at every prefix length: the planner never invents an error and reports the same
    Is_End_Of_Frame as the oracle fed the same prefix."""
    data = _read("concat390.lz4")
    cuts = list(range(0, len(data), 7)) + [len(data)]
    streams = [data[:c] for c in cuts]
    b, _ = _plan(streams)
    for k, c in enumerate(cuts):
        exc, out, eof, msg = oracle.decode_stream(streams[k], chunk=0, out_cap=1 << 16)
        ho = b.host_outcome(k)
        assert exc == "OK"
        assert ho["exception"] == "OK" and ho["end_of_frame"] == eof, (c, ho, eof)


def test_skippable_then_large_block_frame_quirk(oracle):
    """lib/lz4ada.adb:177 pins the reservation to 64 KiB after a skippable frame; a following
    4 MiB-block frame raises Too_Little_Memory in the reference.  Host layer = oracle."""
    stream = _read("skippable.lz4") + _read("z9m.lz4")
    exc, out, eof, msg = oracle.decode_stream(stream, chunk=0, out_cap=1 << 16)
    assert exc == "TOO_LITTLE_MEMORY"
    b, _ = _plan([stream])
    ho = b.host_outcome(0)
    assert ho["exception"] == exc and ho["message"] == msg


def test_gpu_paths_fail_loudly_without_device():
    """No CPU fallback: without a CUDA device a block decode is a Device_Error, not a result."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lz.Device_Error):
        lz.DeviceContext(0)
    d = lz.Init()
    data = _read("z100.lz4")
    with pytest.raises(lz.Device_Error):
        pos = 0
        while pos < len(data):
            c, out, of, ol = d.Update(data[pos:])
            pos += c


def test_host_tools(oracle):
    """tools/lz4ada_tools.py: xxhash32 and hdrinfo counterparts (host layer only, no GPU)."""
    import subprocess
    import sys
    tool = os.path.join(ROOT, "tools", "lz4ada_tools.py")
    data = _read("t100k.bin")
    out = subprocess.run([sys.executable, tool, "xxhash32"], input=data, capture_output=True)
    assert out.stdout.decode().strip() == "%08x" % oracle.xxh32(data)
    out = subprocess.run([sys.executable, tool, "hdrinfo"], input=_read("t301k.lz4"), capture_output=True).stdout.decode()
    assert "block max        262144" in out and "block checksum   True" in out and "block independ.  False" in out
    out = subprocess.run([sys.executable, tool, "hdrinfo"], input=_read("corruptedmagic.err"), capture_output=True).stdout.decode()
    assert out.strip() == MAN["error"]["corruptedmagic"]["eds"]


def test_planner_places_few_big_blocks_as_chains_of_one():
    """Big independent blocks (>= 64 KiB of compressed bytes) become chains of one (LZ4B200_BLK_SOLO | CHAINED |
    FIRST_OF_FRAME) for the chain kernel while a batch holds few of them, and stay ordinary K1 blocks beyond the
    limit (csrc/host/batch.cpp; LZ4B200_SOLO_MAX moves the limit, read once per process: a process each)."""
    import subprocess
    import sys
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import bo_lz4_ada_b200 as lz
from tools import corpus
import test_abi_host_cpu as T
text = corpus.text_like(700000, seed=3)
frames = [corpus.build_frame(text[i * 1000:i * 1000 + 600000], 5, False, True) for i in range(3)]   # 256 KiB blocks: 2 big + 1 small each
frames.append(corpus.build_frame(text[:200000], 4, False, True))                                       # 64 KiB blocks: never
frames.append(corpus.build_frame(bytes(600000), 6, False, True))                                       # zeros: a few hundred compressed bytes
b, _ = T._plan(frames)
flags = [b.block_desc(i).flags for i in range(b.block_count)]
sizes = [b.block_desc(i).src_len for i in range(b.block_count)]
big = [i for i, n in enumerate(sizes) if n >= 65536]
assert len(big) == 6, sizes
solo = [i for i, f in enumerate(flags) if f & 32]
print("solo", len(solo), "big", len(big))
want = big if int(os.environ.get("LZ4B200_SOLO_MAX", "400")) >= len(big) and os.environ.get("LZ4B200_SOLO", "") != "0" else []
assert solo == want, (solo, want)
for i in solo:
    assert flags[i] & 8 and flags[i] & 16            # CHAINED, FIRST_OF_FRAME: a chain of its own
for i in range(len(flags)):
    if i not in solo:
        assert not flags[i] & 8                       # independent frames: nothing else is chained
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env_extra, expect in (({}, "solo 6 big 6"), ({"LZ4B200_SOLO_MAX": "5"}, "solo 0 big 6"), ({"LZ4B200_SOLO": "0"}, "solo 0 big 6")):
        env = dict(os.environ)
        env.pop("LZ4B200_SOLO", None)
        env.pop("LZ4B200_SOLO_MAX", None)
        env.update(env_extra)
        p = subprocess.run([sys.executable, "-c", code % (root, root)], capture_output=True, text=True, env=env, timeout=120)
        assert p.returncode == 0 and expect in p.stdout, (env_extra, p.stdout[-300:], p.stderr[-1500:])
