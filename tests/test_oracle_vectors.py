"""Pins the CPU oracle against every golden vector the reference's own test-suite holds
(/root/reference/test_suite/lz4test.adb).  CPU-only; runs everywhere."""
import hashlib
import json
import os

import pytest

import oracle_binding
from oracle_binding import OracleError

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
INLINE = json.load(open(os.path.join(GOLDEN, "inline_cases.json")))
GOOD = sorted(MAN["good"])
ERR = sorted(MAN["error"])
# 1-byte feeding of the multi-megabyte vectors is slow through ctypes; the C driver does the loop
# natively so every vector is still covered at both granularities.


def _read(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def _check_output(stem, out):
    e = MAN["good"][stem]
    assert len(out) == e["size"]
    assert hashlib.sha256(out).hexdigest() == e["sha256"]
    if e["bin_in_tree"]:
        assert out == _read(stem + ".bin")


@pytest.mark.parametrize("stem", GOOD)
@pytest.mark.parametrize("chunk", [4096, 1])
def test_good_case(oracle, stem, chunk):
    """Test_Good_Case_4K / _1B, lz4test.adb:250-270: Init(For_All) + Update, EOF /= No at the end."""
    exc, out, eof, msg = oracle.decode_stream(_read(stem + ".lz4"), chunk=chunk,
                                              out_cap=MAN["good"][stem]["size"] + 64)
    assert exc == "OK", msg
    assert eof != "No"
    _check_output(stem, out)


@pytest.mark.parametrize("stem", GOOD)
def test_good_case_whole_input(oracle, stem):
    exc, out, eof, msg = oracle.decode_stream(_read(stem + ".lz4"), chunk=0,
                                              out_cap=MAN["good"][stem]["size"] + 64)
    assert exc == "OK", msg
    assert eof != "No"
    _check_output(stem, out)


@pytest.mark.parametrize("stem", ERR)
def test_error_case(oracle, stem):
    """Test_Error_Case, lz4test.adb:280-351: first 10 001 bytes, Init_With_Header(Single_Frame),
    message must equal the first line of the .eds file exactly."""
    data = _read(stem + ".err")[:10001]
    exc, out, msg = oracle.decode_error_case(data)
    assert exc != "OK"
    assert msg == MAN["error"][stem]["eds"]
    assert msg.startswith("raised LZ4ADA." + exc + " : ")


def test_xxh32_kat_individual_bytes(oracle):
    """Test_Good_Hash_Individual_Bytes, lz4test.adb:129-147."""
    case = INLINE["xxh32_kat"]
    h = oracle.hasher()
    for b in bytes.fromhex(case["input_hex"]):
        h.update(bytes([b]))
    assert "%08x" % h.final() == case["expect"]
    assert "%08x" % oracle.xxh32(bytes.fromhex(case["input_hex"])) == case["expect"]


def test_decompress_individual_bytes(oracle):
    """Test_Good_Decompress_Individual_Bytes, lz4test.adb:149-214 (Init_With_Header(For_All), bytewise,
    'consumed 0 => produced output')."""
    case = INLINE["two_legacy_frames"]
    tc = bytes.fromhex(case["input_hex"])
    ctx, consumed0 = oracle.init_with_header(tc, "For_All")
    have = b""
    for i in range(consumed0, len(tc)):
        consumed = 0
        while consumed == 0:
            consumed, out, of, ol = ctx.update(tc[i:i + 1])
            assert not (consumed == 0 and ol < of), "no output produced but expected"
            have += out
    assert have == bytes.fromhex(case["expect_hex"])


def test_hello_block(oracle):
    """Test_Good_Hello_Block, lz4test.adb:216-248."""
    case = INLINE["hello_block"]
    tc = bytes.fromhex(case["input_hex"])
    ctx = oracle.init_for_block(len(tc))
    consumed, out, of, ol = ctx.update(tc)
    assert consumed == len(tc)
    assert ctx.is_end_of_frame() == "Yes"
    assert out == bytes.fromhex(case["expect_hex"])
    assert of == 0


def test_reservation_exceeded(oracle):
    """Test_Error_Case_Reservation_Exceeded, lz4test.adb:353-382."""
    case = INLINE["reservation_exceeded"]
    with pytest.raises(OracleError) as ei:
        oracle.init_with_header(bytes.fromhex(case["input_hex"]), "SZ_64_KiB")
    assert ei.value.name == "TOO_LITTLE_MEMORY"
    assert ei.value.message == ("raised LZ4ADA.TOO_LITTLE_MEMORY : LZ4 header requres reservation SZ_1_MIB, "
                                "but API call requested that only SZ_64_KIB be used. This frame cannot be "
                                "processed under the given constraints.")


def test_unexpected_multi_frame(oracle):
    """Test_Error_Case_Unexpected_Multi_Frame, lz4test.adb:384-430."""
    tc = bytes.fromhex(INLINE["unexpected_multi_frame"]["input_hex"])
    ctx, total = oracle.init_with_header(tc, "Single_Frame")
    with pytest.raises(OracleError) as ei:
        while total < len(tc):
            consumed, out, of, ol = ctx.update(tc[total:])
            total += consumed
    assert ei.value.name == "DATA_CORRUPTION"
    assert "looks like the beginning of another frame" in ei.value.message


def test_eof_tristate(oracle):
    """Is_End_Of_Frame, lib/lz4ada.adb:906-915: legacy is always Maybe, modern No -> Yes."""
    ctx = oracle.init()
    data = _read("z100legacy.lz4")
    pos = 0
    while pos < len(data):
        c, out, _, _ = ctx.update(data[pos:])
        pos += c
    assert ctx.is_end_of_frame() == "Maybe"
    ctx = oracle.init()
    data = _read("z100.lz4")
    pos = 0
    seen = set()
    while pos < len(data):
        c, out, _, _ = ctx.update(data[pos:])
        pos += c
        seen.add(ctx.is_end_of_frame())
    assert ctx.is_end_of_frame() == "Yes" and "No" in seen


def test_against_liblz4_and_xxhash(oracle):
    """Independent cross-check: liblz4 1.9.4 frames of random-ish data, python xxhash."""
    import ctypes
    import numpy as np
    xxhash = pytest.importorskip("xxhash")
    try:
        lz4 = ctypes.CDLL("liblz4.so.1")
    except OSError:
        pytest.skip("liblz4.so.1 not loadable")
    rng = np.random.default_rng(7)
    for n in [0, 1, 15, 16, 17, 100, 4095, 65536, 200000]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.xxh32(data) == xxhash.xxh32(data, seed=0).intdigest()
    lz4.LZ4F_compressFrameBound.restype = ctypes.c_size_t
    lz4.LZ4F_compressFrameBound.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    lz4.LZ4F_compressFrame.restype = ctypes.c_size_t
    lz4.LZ4F_compressFrame.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                       ctypes.c_void_p]
    words = [bytes(rng.integers(97, 123, rng.integers(2, 10), dtype=np.uint8)) for _ in range(500)]
    text = b" ".join(words[i] for i in rng.zipf(1.3, 60000) % 500)
    for data in [text, bytes(300000), text[:70000] + bytes(1000) + text[:5000]]:
        cap = lz4.LZ4F_compressFrameBound(len(data), None)
        dst = ctypes.create_string_buffer(cap)
        n = lz4.LZ4F_compressFrame(dst, cap, data, len(data), None)
        frame = dst.raw[:n]
        exc, out, eof, msg = oracle.decode_stream(frame, chunk=4096, out_cap=len(data) + 64)
        assert exc == "OK", msg
        assert out == data and eof == "Yes"
