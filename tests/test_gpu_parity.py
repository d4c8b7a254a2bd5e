"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C-ABI,
against the CPU oracle and the reference's golden vectors.  Bit-exact: this is byte / integer work.
Nothing here reads /root/reference."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

import bo_lz4_ada_b200 as lz
from tools import corpus

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
INLINE = json.load(open(os.path.join(GOLDEN, "inline_cases.json")))
GOOD = sorted(MAN["good"])
ERR = sorted(MAN["error"])


def _read(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def ctx():
    c = lz.DeviceContext(0)
    c.make_default()
    yield c
    lz.lib().lz4ada_set_device_context(None)
    c.close()


def _check_output(stem, out):
    e = MAN["good"][stem]
    assert len(out) == e["size"]
    assert hashlib.sha256(out).hexdigest() == e["sha256"]
    if e["bin_in_tree"]:
        assert out == _read(stem + ".bin")


def _drive(dec, data, chunk):
    """The feeding loop of Test_Good_Case_Inner (test_suite/lz4test.adb:32-83)."""
    out = bytearray()
    pos = 0
    while pos < len(data):
        end = min(len(data), pos + chunk) if chunk else len(data)
        while pos < end:
            c, o, of, ol = dec.Update(data[pos:end])
            out += o
            pos += c
            assert c > 0 or len(o) > 0, "no progress"
    return bytes(out)


# ------------------------------------------------------------------ streaming API (Update)

@pytest.mark.parametrize("stem", GOOD)
def test_good_case_4k(ctx, stem):
    """Test_Good_Case_4K (lz4test.adb:254-267) through lz4ada_update: every block decoded on the GPU."""
    before = ctx.launch_count()
    dec = lz.Init()
    out = _drive(dec, _read(stem + ".lz4"), 4096)
    assert dec.Is_End_Of_Frame() != "No"
    _check_output(stem, out)
    if MAN["good"][stem]["size"] > 0:
        assert ctx.launch_count() > before, "no kernel launched: not the CUDA path"


@pytest.mark.parametrize("stem", [s for s in GOOD if MAN["good"][s]["lz4"]["size"] <= 300000])
def test_good_case_1b(ctx, stem):
    """Test_Good_Case_1B (lz4test.adb:252, 259-270): one byte per call."""
    dec = lz.Init()
    out = _drive(dec, _read(stem + ".lz4"), 1)
    assert dec.Is_End_Of_Frame() != "No"
    _check_output(stem, out)


def test_update_step_parity_with_oracle(ctx, oracle):
    """Per-call parity: Num_Consumed, Output_First/Last, the bytes and Is_End_Of_Frame of every
    Update call equal the oracle's, for awkward chunk sizes."""
    for stem in ["t300k", "t301k", "concat390", "z101legacyplus", "skipz100", "z2841", "concatlegacy"]:
        data = _read(stem + ".lz4")
        for chunk in (1 if len(data) < 2000 else 997, 65536 + 13):
            a, b = lz.Init(), oracle.init()
            pos = 0
            while pos < len(data):
                piece = data[pos:pos + chunk]
                ra = a.Update(piece)
                rb = b.update(piece)
                assert ra == rb, (stem, chunk, pos)
                assert a.Is_End_Of_Frame() == b.is_end_of_frame()
                pos += ra[0]



def _step_parity(lz_dec, o_dec, data, chunk):
    """Feed both decompressors the same pieces; every call must agree on Num_Consumed, the bytes,
    Output_First/Last and Is_End_Of_Frame, and on the exception that ends the run (if any)."""
    pos = 0
    while pos < len(data):
        piece = data[pos:pos + chunk]
        ea = eb = None
        try:
            ra = lz_dec.Update(piece)
        except lz.LZ4AdaError as e:
            ea = (e.ada_name, e.information)
        try:
            rb = o_dec.update(piece)
        except Exception as e:   # OracleError
            eb = (e.name, "raised LZ4ADA.%s : %s" % (e.name, e) if False else None)
            eb = (e.name, None)
        if ea or eb:
            assert ea is not None and eb is not None and ea[0] == eb[0], (pos, ea, eb)
            return ea
        assert ra == rb, (chunk, pos, ra[0], rb[0], ra[2:], rb[2:], len(ra[1]), len(rb[1]))
        assert lz_dec.Is_End_Of_Frame() == o_dec.is_end_of_frame()
        assert ra[0] > 0 or ra[1], "no progress"
        pos += ra[0]
    return None


def test_update_read_ahead_step_parity(ctx, oracle):
    """Large Inputs trigger the read-ahead (many blocks decoded by one K1 launch, handed out one per Update
    call): per-call parity with the oracle must be exactly what it is for small Inputs -- good vectors, synthetic
    frames of every block size with and without block checksums / content size, concatenations, and corrupted
    streams (the exception must fire in the same call)."""
    cases = []
    for stem in ["t300k", "t301k", "t1111k", "b3444k", "concat390", "z101legacyplus", "skipz100", "z2841", "concatlegacy", "z9m"]:
        cases.append((stem, _read(stem + ".lz4")))
    text = corpus.text_like(900000, seed=21)
    rle = corpus.rle_like(300000, seed=22)
    rnd = corpus.random_bytes(150000, seed=23)
    mix = text[:400000] + rnd + rle + text[400000:]
    for code in (4, 5, 6, 7):
        for bchk in (False, True):
            cases.append(("mix-b%d-%d" % (code, bchk), corpus.build_frame(mix, code, bchk, True, True)))
    cases.append(("mix-linked", corpus.build_frame(mix, 4, True, True, True, independent=False)))
    cases.append(("short-interior", corpus.build_frame(text[:300000], 4, True, True, block_size=50000)))
    cases.append(("two-frames", corpus.build_frame(text[:200000], 4, True, True) + corpus.build_frame(rle[:99999], 5, False, True, True)))
    rng = np.random.default_rng(31)
    base = corpus.build_frame(mix[:500000], 4, True, True, True)
    for k, bad in enumerate(_mutations(base, rng, 24)):
        cases.append(("mutation-%d" % k, bad))
    ends = {}
    for name, data in cases:
        for chunk in (len(data), 1 << 20, 200000 + 7):
            e = _step_parity(lz.Init(), oracle.init(), data, chunk)
            ends.setdefault(name, e)
            assert ends[name] == e or (ends[name] and e and ends[name][0] == e[0]), (name, chunk, ends[name], e)
    assert any(v for k, v in ends.items() if k.startswith("mutation")), "no mutation raised"


def test_update_read_ahead_changed_input(ctx):
    """The caller does not have to present the same bytes again: if what follows differs from what was decoded
    ahead, the staged blocks are dropped and the new bytes are decoded."""
    a = corpus.text_like(400000, seed=41)
    b = corpus.text_like(400000, seed=42)
    fa, fb = corpus.build_frame(a, 4, False, False), corpus.build_frame(b, 4, False, False)
    # same 7-byte header, different blocks: take the first block of `fa` out of the whole of `fa` (the blocks behind
    # it are decoded ahead), then continue with the second block of `fb`
    assert fa[:7] == fb[:7]
    dec = lz.Init()
    pos, o1 = 0, b""
    while not o1:
        c, o1, _, _ = dec.Update(fa[pos:])
        pos += c
    assert o1 == a[:65536]
    n1a = struct.unpack_from("<I", fa, 7)[0] & 0x7fffffff
    n1b = struct.unpack_from("<I", fb, 7)[0] & 0x7fffffff
    assert pos == 7 + 4 + n1a
    pos_b = 7 + 4 + n1b
    out = bytearray(o1)
    while pos_b < len(fb):
        c, o, _, _ = dec.Update(fb[pos_b:])
        assert c > 0 or o
        out += o
        pos_b += c
    assert bytes(out) == a[:65536] + b[65536:]

@pytest.mark.parametrize("stem", ERR)
def test_error_case(ctx, stem):
    """Test_Error_Case (lz4test.adb:280-351): first 10 001 bytes, Init_With_Header(Single_Frame);
    Exception_Information must equal the .eds line exactly."""
    data = _read(stem + ".err")[:10001]
    with pytest.raises(lz.LZ4AdaError) as ei:
        dec, total = lz.Init_With_Header(data, "Single_Frame")
        while total < len(data):
            c, o, of, ol = dec.Update(data[total:])
            assert c > 0, "No more data accepted but no exception signalled"
            total += c
    assert ei.value.information == MAN["error"][stem]["eds"]
    assert not isinstance(ei.value, (lz.Device_Error, lz.Constraint_Error, lz.Assertion_Error))


def test_decompress_individual_bytes(ctx):
    """Test_Good_Decompress_Individual_Bytes (lz4test.adb:149-214)."""
    case = INLINE["two_legacy_frames"]
    tc = bytes.fromhex(case["input_hex"])
    dec, consumed0 = lz.Init_With_Header(tc, "For_All")
    have = b""
    for i in range(consumed0, len(tc)):
        consumed = 0
        while consumed == 0:
            consumed, out, of, ol = dec.Update(tc[i:i + 1])
            assert not (consumed == 0 and ol < of)
            have += out
    assert have == bytes.fromhex(case["expect_hex"])


def test_hello_block(ctx):
    """Test_Good_Hello_Block (lz4test.adb:216-248)."""
    case = INLINE["hello_block"]
    tc = bytes.fromhex(case["input_hex"])
    dec = lz.Init_For_Block(len(tc))
    consumed, out, of, ol = dec.Update(tc)
    assert consumed == len(tc) and dec.Is_End_Of_Frame() == "Yes"
    assert out == bytes.fromhex(case["expect_hex"]) and of == 0


def test_hello_block_in_pieces(ctx):
    """Raw-block API fed in several chunks (the reference drops 4 cached bytes here, Appendix C)."""
    tc = bytes.fromhex(INLINE["hello_block"]["input_hex"])
    dec = lz.Init_For_Block(len(tc))
    out = b""
    for i in range(0, len(tc), 5):
        c, o, _, _ = dec.Update(tc[i:i + 5])
        assert c == len(tc[i:i + 5])
        out += o
    assert out == b"Hello, world."


def test_unexpected_multi_frame(ctx):
    """Test_Error_Case_Unexpected_Multi_Frame (lz4test.adb:384-430)."""
    tc = bytes.fromhex(INLINE["unexpected_multi_frame"]["input_hex"])
    dec, total = lz.Init_With_Header(tc, "Single_Frame")
    with pytest.raises(lz.Data_Corruption) as ei:
        while total < len(tc):
            c, o, of, ol = dec.Update(tc[total:])
            total += c
    assert "looks like the beginning of another frame" in str(ei.value)


# ------------------------------------------------------------------ batched entry point

def test_batch_all_good_vectors(ctx):
    """All 24 good vectors as one batch: K1 (independent), K4 (linked: t300k/t301k/z2841/b3444k),
    K5 pre-sizing (concatenated frames), K3 content checksums."""
    streams = [_read(s + ".lz4") for s in GOOD]
    before = ctx.launch_count()
    res = lz.batch_decompress(ctx, streams)
    assert ctx.launch_count() - before >= 3
    for stem, (exc, out, eof, msg) in zip(GOOD, res):
        assert exc == "OK", (stem, msg)
        assert eof != "No", stem
        _check_output(stem, out)


@pytest.mark.parametrize("stem", ["empty", "emptycraft", "skippable", "z1", "t2"])
def test_batch_of_one_tiny_stream(ctx, stem):
    """A batch that is nothing but one frame without blocks (or one skippable frame, or one stored byte): no kernel has
    anything to do for some of them, the outcome still comes from the fold (content checksum of no bytes included)."""
    (exc, out, eof, msg), = lz.batch_decompress(ctx, [_read(stem + ".lz4")])
    assert exc == "OK", msg
    assert eof != "No"
    _check_output(stem, out)


def test_batch_error_vectors_match_oracle(ctx, oracle):
    """Every .err vector through the batch path (Init(For_All) semantics) = the oracle's
    Init(For_All)+Update loop: same exception, same text, same bytes before the error."""
    streams = [_read(s + ".err") for s in ERR]
    res = lz.batch_decompress(ctx, streams)
    for stem, data, (exc, out, eof, msg) in zip(ERR, streams, res):
        oexc, oout, oeof, omsg = oracle.decode_stream(data, chunk=0, out_cap=1 << 22)
        assert (exc, msg) == (oexc, omsg), stem
        assert out == oout, stem



def test_batch_block_outgrows_block_max(ctx, oracle):
    """The reference bounds a block's output by its caller's Buffer only (lib/lz4ada.adb:54, 813-820): under
    Init(For_All) it accepts cntblkszoverflow.err, a 64 KiB-block frame whose single block inflates to 100 KiB.
    The batch call decodes such a stream again as a chain under that bound, moved behind the planned output
    when it outgrows its region; without spare room it ends with the oracle-style Output-buffer text."""
    data = _read("cntblkszoverflow.err")
    oexc, oout, oeof, omsg = oracle.decode_stream(data, chunk=0, out_cap=1 << 22)
    assert oexc == "OK" and len(oout) == 102400
    text = corpus.text_like(200000, seed=5)
    good = corpus.build_frame(text, 4, True, True)
    info = {}
    res = lz.batch_decompress(ctx, [good, data, good, data], info=info)
    assert info["retried_streams"] == 2
    for k, plain in ((0, text), (1, oout), (2, text), (3, oout)):
        assert res[k][0] == "OK" and res[k][1] == plain, (k, res[k][0], res[k][3])
    # no spare capacity: the overflow is reported against the region the stream has (one block maximum)
    b = lz.Batch(ctx, data, [(0, len(data))])
    d_src, d_dst = ctx.alloc(len(data) + 64), ctx.alloc(b.output_bytes + 64)
    b.upload(d_src)
    b.run(d_src, d_dst)
    r = b.results()[0]
    assert r["exception"] == "DATA_CORRUPTION" and r["out_len"] == 0
    assert r["message"] == ("raised LZ4ADA.DATA_CORRUPTION : Output buffer exhausted. Decompressed data does not fit "
                            "into the 65536 bytes provided.")
    b.close()
    ctx.free(d_src)
    ctx.free(d_dst)
    # a frame of several blocks, one of which is 30 KiB of compressed zeros-with-noise inflating past 64 KiB:
    # build it by hand -- a 4 MiB-block payload inside a frame that declares 64 KiB blocks
    big = corpus.compress_block(bytes(150000) + text[:20000])
    hdr = corpus.frame_header(4, False, True)
    frame = hdr + struct.pack("<I", len(big)) + big + struct.pack("<I", 0) + struct.pack("<I", corpus.xxh32(bytes(150000) + text[:20000]))
    oexc, oout, oeof, omsg = oracle.decode_stream(frame, chunk=0, out_cap=1 << 22)
    assert oexc == "OK" and oout == bytes(150000) + text[:20000]
    (exc, out, eof, msg), = lz.batch_decompress(ctx, [frame])
    assert (exc, out) == (oexc, oout), msg


@pytest.mark.parametrize("reservation", ["Single_Frame", "Use_First"])
def test_batch_error_vectors_init_with_header(ctx, oracle, reservation):
    """Test_Error_Case (lz4test.adb:280-351) through the batch call: first 10 001 bytes of every .err vector,
    decoded as Init_With_Header(all, Single_Frame) + Update would -- the message must be the .eds line itself."""
    streams = [_read(s + ".err")[:10001] for s in ERR] + [_read(s + ".lz4") for s in GOOD]
    res = lz.batch_decompress(ctx, streams, Reservation=reservation)
    for stem, data, (exc, out, eof, msg) in zip(ERR, streams, res):
        if reservation == "Single_Frame":
            assert msg == MAN["error"][stem]["eds"], stem
            oexc, oout, omsg = oracle.decode_error_case(data)
            assert (exc, out) == (oexc, oout), stem
        elif stem != "trailingbytes":   # the one vector whose error is the Single_Frame policing itself
            assert msg == MAN["error"][stem]["eds"], stem
    for stem, data, (exc, out, eof, msg) in zip(GOOD, streams[len(ERR):], res[len(ERR):]):
        oexc, oout, omsg = _oracle_with_header(oracle, data, reservation)
        assert (exc, msg, out) == (oexc, omsg, oout), stem
        if exc == "OK":
            _check_output(stem, out)


def _oracle_with_header(oracle, data, reservation):
    """Init_With_Header(data, reservation) + Update until the input is used up -> (exception, output, message)."""
    import oracle_binding
    out = bytearray()
    try:
        o, pos = oracle.init_with_header(data, reservation)
        while pos < len(data):
            c, piece, _, _ = o.update(data[pos:])
            out += piece
            pos += c
            assert c > 0 or piece
    except oracle_binding.OracleError as e:
        return e.name, bytes(out), e.message
    return "OK", bytes(out), ""


def test_context_lifetime(oracle):
    """A device context is reference counted: closing it before the objects made on it (the order the driver's
    smoke() used, and the order an Ada finaliser may pick) must neither crash nor stop them working; a second
    context made afterwards starts clean (no stale scratch keyed by address)."""
    import ctypes
    import gc
    data = _read("t300k.lz4")
    plain = _read("t300k.bin")
    c1 = lz.DeviceContext(0)
    c1.make_default()
    dec = lz.Init()
    c0, o0, _, _ = dec.Update(data[:70000])          # the stream object now exists on c1
    lz.lib().lz4ada_set_device_context(None)
    c1.close()                                        # creator's reference gone; dec still holds one
    got, pos = bytearray(o0), c0
    while pos < len(data):
        c, o, _, _ = dec.Update(data[pos:pos + 4096])
        got += o
        pos += c
    assert bytes(got) == plain
    dec.close()
    del dec
    gc.collect()
    # one-shot calls on contexts that come and go (the scratch pool is owned by its context)
    text = corpus.text_like(400000, seed=9)
    frame = corpus.build_frame(text, 4, True, True)
    for rep in range(12):
        c = lz.DeviceContext(0)
        items = (lz.BatchItem * 1)()
        items[0].src_off, items[0].src_len = 0, len(frame)
        results = (lz.BatchResult * 1)()
        out = bytearray(len(text) + (1 << 16))
        addr = (ctypes.c_uint8 * len(out)).from_buffer(out)
        rc = lz.lib().lz4ada_batch_decompress(c.handle, frame, len(frame), addr, len(out), 1, items,
                                              lz.RESERVATIONS["For_All"], results, None, 0)
        assert rc == 0 and results[0].exception == 0
        assert bytes(out[results[0].dst_off:results[0].dst_off + results[0].out_len]) == text
        c.close()
    # a batch outliving its context
    c = lz.DeviceContext(0)
    b = lz.Batch(c, frame, [(0, len(frame))])
    d_src, d_dst = c.alloc(len(frame) + 64), c.alloc(b.output_bytes + 64)
    b.upload(d_src)
    handle = c.handle
    lz.lib().lz4b200_retain(handle)   # keep our own reference for the device buffers below
    c.close()
    b.run(d_src, d_dst)
    assert b.results()[0]["exception"] == "OK"
    b.close()
    c.handle = handle
    assert c.d2h(d_dst, len(text)) == text
    c.free(d_src)
    c.free(d_dst)
    c.close()


def test_batch_run_twice_with_retries_and_chains(ctx):
    """lz4ada_batch_run is repeatable: a run that had to decode some streams again as chains (short interior
    blocks) must leave the plan's own chain table (linked frames) intact for the next run -- every run into a
    cleared output buffer has to produce all the bytes again."""
    text = corpus.text_like(600000, seed=31)
    streams = [corpus.build_frame(text[:300000], 4, True, True, independent=False),           # linked: a plan chain
               corpus.build_frame(text[100000:400000], 4, False, True, block_size=30000),     # short blocks: retried
               corpus.build_frame(text[200000:], 5, True, True, independent=False),           # linked
               corpus.build_frame(text[:200000], 4, True, True)]
    plains = [text[:300000], text[100000:400000], text[200000:], text[:200000]]
    src = b"".join(streams)
    offs, pos = [], 0
    for st in streams:
        offs.append((pos, len(st)))
        pos += len(st)
    b = lz.Batch(ctx, src, offs)
    need = b.output_bytes
    d_src, d_dst = ctx.alloc(len(src) + 64), ctx.alloc(need + 64)
    b.upload(d_src)
    for rep in range(3):
        lz.lib().lz4b200_memset(ctx.handle, d_dst, 0xA5 + rep, need)
        b.run(d_src, d_dst)
        assert b.retried_streams() == 1
        for plain, r in zip(plains, b.results()):
            assert r["exception"] == "OK", (rep, r["message"])
            assert ctx.d2h(d_dst + r["dst_off"], r["out_len"]) == plain, rep
    b.close()
    ctx.free(d_src)
    ctx.free(d_dst)


def test_batch_exact_sizing(ctx, oracle):
    """Exact sizing (K5 over every block, blocks placed back to back): same outcomes as the default placement for
    good vectors, synthetic frames and corrupted streams; frames with short interior blocks no longer go through
    the decode-again-as-a-chain retry."""
    streams = [_read(s + ".lz4") for s in GOOD]
    for stem, (exc, out, eof, msg) in zip(GOOD, lz.batch_decompress(ctx, streams, exact_sizing=True)):
        assert exc == "OK", (stem, msg)
        _check_output(stem, out)
    cases = _synthetic_frames()
    frames = [c[1] for c in cases]
    i0, i1 = {}, {}
    plain_res = lz.batch_decompress(ctx, frames, info=i0)
    exact_res = lz.batch_decompress(ctx, frames, exact_sizing=True, info=i1)
    assert plain_res == exact_res
    for (name, frame, plain), (exc, out, eof, msg) in zip(cases, exact_res):
        assert exc == "OK" and out == plain, (name, msg)
    # the short-interior-blocks case needs the retry by default, never with exact sizing; linked frames that reach
    # into the previous block are chains either way
    text = corpus.text_like(300000, seed=3)
    short = [corpus.build_frame(text, 4, False, True, block_size=50000), corpus.build_frame(text[:170000], 4, True, True, block_size=30000)]
    j0, j1 = {}, {}
    r0 = lz.batch_decompress(ctx, short, info=j0)
    r1 = lz.batch_decompress(ctx, short, exact_sizing=True, info=j1)
    assert r0 == r1 and r1[0][1] == text and r1[1][1] == text[:170000]
    assert j0["retried_streams"] == 2 and j1["retried_streams"] == 0, (j0, j1)
    # corrupted streams: the same exception, message and bytes-before-error as the oracle
    rng = np.random.default_rng(17)
    mix = text[:60000] + bytes(5000) + corpus.random_bytes(3000, seed=1) + text[60000:120000]
    bad = _mutations(corpus.build_frame(mix, 4, True, True, True), rng, 40) + _mutations(short[0], rng, 20)
    for k, (data, (exc, out, eof, msg)) in enumerate(zip(bad, lz.batch_decompress(ctx, bad, exact_sizing=True))):
        oexc, oout, oeof, omsg = oracle.decode_stream(data, chunk=0, out_cap=1 << 21)
        assert (exc, msg, out) == (oexc, omsg, oout), (k, exc, msg, oexc, omsg)

@pytest.mark.parametrize("g", [-1, 1, 2, 4, 8, 16, 64, 40, 41, 48, 50, 60, 61])
def test_batch_k1_variants(ctx, oracle, g):
    """Every K1 variant (v1 one-warp-per-block, v2 with 1/2/4/8/16 blocks per warp, 64 = v3 CTA per block,
    40..48 = v4 warp per block with a shared-memory ring, 50 = v5 lane per block): good vectors,
    synthetic frames and corrupted streams must all come out exactly as the oracle says."""
    ctx.set_tuning(g)
    try:
        streams = [_read(s + ".lz4") for s in GOOD]
        for stem, (exc, out, eof, msg) in zip(GOOD, lz.batch_decompress(ctx, streams)):
            assert exc == "OK", (stem, msg)
            _check_output(stem, out)
        cases = _synthetic_frames()
        for (name, frame, plain), (exc, out, eof, msg) in zip(cases, lz.batch_decompress(ctx, [c[1] for c in cases])):
            assert exc == "OK" and out == plain, (name, msg)
        rng = np.random.default_rng(100 + g)
        text = corpus.text_like(60000, seed=19)
        mix = text[:20000] + bytes(5000) + corpus.random_bytes(3000, seed=1) + text[20000:45000]
        bases = [corpus.build_frame(mix, 4, False, True, True), corpus.build_frame(mix, 4, True, True, False),
                 corpus.build_frame(text, 4, False, False, False, block_size=7000)]
        bad = []
        for base in bases:
            bad += _mutations(base, rng, 40)
        for k, (data, (exc, out, eof, msg)) in enumerate(zip(bad, lz.batch_decompress(ctx, bad))):
            oexc, oout, oeof, omsg = oracle.decode_stream(data, chunk=0, out_cap=1 << 21)
            assert (exc, msg, out) == (oexc, omsg, oout), (g, k, exc, msg, oexc, omsg)
    finally:
        ctx.set_tuning(0)


def _synthetic_frames():
    rng = np.random.default_rng(11)
    text = corpus.text_like(700000, seed=3)
    rle = corpus.rle_like(600000, seed=4)
    rnd = corpus.random_bytes(200000, seed=5)
    mix = text[:150000] + rnd[:70000] + rle[:150000] + text[150000:300000]
    cases = []
    for code in (4, 5, 6, 7):
        for indep in (True, False):
            for bchk in (True, False):
                cases.append(("mix-b%d-%s-%s" % (code, "i" if indep else "l", "bc" if bchk else "nb"),
                              corpus.build_frame(mix, code, bchk, True, True, indep), mix))
    cases.append(("text-64k", corpus.build_frame(text, 4, True, True), text))
    cases.append(("rle-4m", corpus.build_frame(rle, 7, False, True), rle))
    cases.append(("random-stored", corpus.build_frame(rnd, 4, True, True), rnd))
    cases.append(("legacy", corpus.build_legacy_frame(text + rle, block_size=1 << 20), text + rle))
    cases.append(("greedy-enc", corpus.build_frame(text, 5, True, True, encoder="greedy"), text))
    cases.append(("short-interior-blocks", corpus.build_frame(text[:300000], 4, False, True, block_size=50000),
                  text[:300000]))
    cases.append(("linked-short-blocks", corpus.build_frame(text[:300000], 4, True, True, independent=False,
                                                            block_size=30000), text[:300000]))
    cases.append(("dict-id-field", corpus.build_frame(text[:5000], 4, False, True, dict_id=7), text[:5000]))
    cases.append(("concat+skip", corpus.skippable_frame(b"x" * 33, 3) + corpus.build_frame(text[:70000], 4) +
                  corpus.build_frame(rle[:99999], 4, True) + corpus.skippable_frame(b"", 15) +
                  corpus.build_frame(b"", 4), text[:70000] + rle[:99999]))
    # tiny inputs and every small period
    for n in (1, 2, 3, 4, 5, 12, 13, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 1000):
        d = bytes(rng.integers(97, 100, n, dtype=np.uint8))
        cases.append(("tiny-%d" % n, corpus.build_frame(d, 4, True, True, True), d))
    for period in list(range(1, 41)) + [63, 64, 65, 100, 511, 512, 513, 4000]:
        pat = bytes(rng.integers(0, 256, period, dtype=np.uint8))
        d = (pat * (70000 // period + 2))[:70000 - period % 7]
        cases.append(("period-%d" % period, corpus.build_frame(d, 4, False, True), d))
    return cases


def test_batch_synthetic_vs_oracle_and_plain(ctx, oracle):
    cases = _synthetic_frames()
    res = lz.batch_decompress(ctx, [c[1] for c in cases])
    for (name, frame, plain), (exc, out, eof, msg) in zip(cases, res):
        assert exc == "OK", (name, msg)
        assert out == plain, name
        oexc, oout, oeof, omsg = oracle.decode_stream(frame, chunk=65536, out_cap=len(plain) + 64)
        assert (oexc, oout, oeof) == ("OK", plain, eof), name


def test_streaming_synthetic(ctx):
    """A few synthetic frames through Update as well (device-resident history across linked blocks)."""
    for name, frame, plain in _synthetic_frames()[:8] + _synthetic_frames()[16:24]:
        dec = lz.Init()
        assert _drive(dec, frame, 100000) == plain, name


def _mutations(frame, rng, n):
    out = []
    for _ in range(n):
        b = bytearray(frame)
        kind = rng.integers(0, 4)
        if kind == 0:      # flip a byte after the header
            i = int(rng.integers(7, len(b)))
            b[i] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:    # overwrite a few bytes
            i = int(rng.integers(7, len(b) - 4))
            b[i:i + 3] = bytes(rng.integers(0, 256, 3, dtype=np.uint8))
        elif kind == 2:    # truncate
            b = b[:int(rng.integers(1, len(b)))]
        else:              # 0xff burst (length-extension stress)
            i = int(rng.integers(7, len(b) - 8))
            b[i:i + 6] = b"\xff" * 6
        out.append(bytes(b))
    return out


def test_fuzz_error_parity_with_oracle(ctx, oracle):
    """Corrupted streams: the batch path and the oracle must report the same first exception with
    the same text and the same output before it (SURVEY.md Appendix A ordering).  Frames without
    block checksums so that the corruption reaches the sequence decoder."""
    rng = np.random.default_rng(2024)
    text = corpus.text_like(40000, seed=9)
    mix = text[:9000] + bytes(3000) + text[9000:14000]
    bases = [corpus.build_frame(mix, 4, False, True, True), corpus.build_frame(mix, 4, True, False, False),
             corpus.build_frame(text, 4, False, False, False, independent=False, block_size=9000),
             corpus.build_legacy_frame(mix), _read("t100k.lz4")[:30000]]
    streams = []
    for base in bases:
        streams += _mutations(base, rng, 60)
    res = lz.batch_decompress(ctx, streams)
    kinds = {}
    for k, (data, (exc, out, eof, msg)) in enumerate(zip(streams, res)):
        oexc, oout, oeof, omsg = oracle.decode_stream(data, chunk=0, out_cap=1 << 21)
        assert (exc, msg) == (oexc, omsg), (k, exc, msg, oexc, omsg)
        assert out == oout, k
        if exc == "OK":
            assert eof == oeof, k
        kinds[msg.split(" : ")[1][:28] if msg else "OK"] = kinds.get(msg[:1], 0) + 1
    assert len(kinds) >= 5, kinds    # the fuzz actually reached several distinct raise sites


def test_fuzz_streaming_error_parity(ctx, oracle):
    """Same idea through Update (Init(For_All), 4 KiB chunks)."""
    rng = np.random.default_rng(77)
    text = corpus.text_like(30000, seed=10)
    base = corpus.build_frame(text, 4, False, True, True, independent=False, block_size=7000)
    for data in _mutations(base, rng, 40):
        oexc, oout, oeof, omsg = oracle.decode_stream(data, chunk=4096, out_cap=1 << 21)
        dec = lz.Init()
        out = bytearray()
        exc, msg = "OK", ""
        try:
            pos = 0
            while pos < len(data):
                end = min(len(data), pos + 4096)
                idle = 0
                while pos < end:
                    c, o, of, ol = dec.Update(data[pos:end])
                    out += o
                    pos += c
                    idle = idle + 1 if (c == 0 and not o) else 0
                    assert idle < 5
        except lz.LZ4AdaError as e:
            exc, msg = e.ada_name, e.information
        assert (exc, msg) == (oexc, omsg)
        assert bytes(out) == oout


# ------------------------------------------------------------------ kernels through the shim

def test_k3_xxh32_spans_alignment_and_lengths(ctx, oracle):
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 1 << 20, dtype=np.uint8).tobytes()
    spans = []
    for n in [0, 1, 3, 4, 15, 16, 17, 31, 32, 33, 47, 48, 127, 128, 129, 1000, 4096, 65536, 300001]:
        for mis in (0, 1, 2, 3, 5, 16):
            spans.append((1000 + mis, n))
    arr = (lz.HashSpan * len(spans))(*[lz.HashSpan(o, n) for o, n in spans])
    d_data = ctx.alloc(len(data))
    d_spans = ctx.alloc(ctypes_sizeof(arr))
    d_out = ctx.alloc(4 * len(spans))
    ctx.h2d(d_data, data)
    ctx.h2d(d_spans, bytes(arr))
    rc = lz.lib().lz4b200_xxh32_spans(ctx.handle, d_data, len(spans), d_spans, d_out)
    assert rc == 0
    got = struct.unpack("<%dI" % len(spans), ctx.d2h(d_out, 4 * len(spans)))
    for (o, n), g in zip(spans, got):
        assert g == oracle.xxh32(data[o:o + n]), (o, n)
    for p in (d_data, d_spans, d_out):
        ctx.free(p)


@pytest.mark.parametrize("n_streams,len_a,len_b", [(8, 70001, 150003), (1200, 9001, 40010), (3600, 5003, 20005), (7200, 3001, 9007)])
def test_k3_frames_ring_variants(ctx, n_streams, len_a, len_b):
    """The content-checksum kernel picks its shared-memory ring by the number of frames (2 KiB groups for a handful
    of frames down to 256-byte groups for thousands).  Every stream is two concatenated frames, so the second frame's
    output starts at an arbitrary byte: misaligned spans that lap each ring several times, checked by the device
    against the checksums in the frames (lib/lz4ada.adb:463-523) -- plus one frame with a wrong checksum."""
    text = corpus.text_like(1 << 20, seed=11)
    variants = []
    for v in range(16):
        a = text[v * 1009:][:len_a + v]
        b = text[5000 + v * 4001:][:len_b + 3 * v]
        variants.append((corpus.build_frame(a, 4, False, True) + corpus.build_frame(b, 4, True, True), a + b))
    streams = [variants[i % 16][0] for i in range(n_streams)]
    bad = bytearray(streams[n_streams // 2])
    bad[-1] ^= 0x40   # the second frame's content checksum
    streams[n_streams // 2] = bytes(bad)
    res = lz.batch_decompress(ctx, streams)
    for i, (exc, out, eof, msg) in enumerate(res):
        if i == n_streams // 2:
            assert exc == "CHECKSUM_ERROR" and "content checksum" in msg, (i, exc, msg)
            continue
        assert exc == "OK", (i, msg)
        assert out == variants[i % 16][1], i


def ctypes_sizeof(a):
    import ctypes
    return ctypes.sizeof(a)


def _raw_block(seqs, last_literals=b""):
    """Hand-assembled LZ4 block from (literals, offset, match_len) triples."""
    out = bytearray()
    for lit, off, ml in seqs:
        ll, mm = len(lit), ml - 4
        out.append((min(ll, 15) << 4) | min(mm, 15))
        if ll >= 15:
            r = ll - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out += lit
        out += struct.pack("<H", off)
        if mm >= 15:
            r = mm - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
    ll = len(last_literals)
    out.append(min(ll, 15) << 4)
    if ll >= 15:
        r = ll - 15
        while r >= 255:
            out.append(255)
            r -= 255
        out.append(r)
    out += last_literals
    return bytes(out)


def _py_decode(block):
    """Tiny pure-Python LZ4 block decoder (test-side cross-check, small cases only)."""
    out = bytearray()
    i = 0
    while i < len(block):
        t = block[i]; i += 1
        ll = t >> 4
        if ll == 15:
            while True:
                b = block[i]; i += 1; ll += b
                if b != 255:
                    break
        out += block[i:i + ll]; i += ll
        if i >= len(block):
            break
        off = block[i] | (block[i + 1] << 8); i += 2
        ml = t & 15
        if ml == 15:
            while True:
                b = block[i]; i += 1; ml += b
                if b != 255:
                    break
        ml += 4
        for _ in range(ml):
            out.append(out[-off])
    return bytes(out)


@pytest.mark.parametrize("g", [-1, 1, 8, 16, 64, 41, 48, 50, 60, 61])
def test_k1_overlap_matrix_direct(ctx, oracle, g):
    """lz4b200_decode_blocks on hand-made blocks: every offset 1..70 x match lengths around the
    warp / vector thresholds, at varying destination alignment (pattern replication, doubling)."""
    rng = np.random.default_rng(8)
    blocks, expect = [], []
    for off in list(range(1, 71)) + [255, 256, 257, 511, 512, 513, 1000]:
        for ml in (4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 200, 513, 2000, 70000):
            lead = bytes(rng.integers(0, 256, max(off, 1) + int(rng.integers(0, 9)), dtype=np.uint8))
            blk = _raw_block([(lead, off, ml)], b"tail!")
            blocks.append(blk)
            exp = bytearray(lead)
            if ml <= 2000:
                for _ in range(ml):
                    exp.append(exp[-off])
            else:
                base = bytes(exp[-off:])
                exp += (base * (ml // off + 2))[:ml]
            expect.append(bytes(exp) + b"tail!")
    cap = 80000
    src = b"".join(blocks)
    descs = (lz.BlkDesc * len(blocks))()
    pos = 0
    for i, blk in enumerate(blocks):
        descs[i].src_off, descs[i].src_len = pos, len(blk)
        descs[i].dst_off, descs[i].dst_cap = i * cap + (i % 16), cap - 16
        descs[i].flags, descs[i].hist_avail = 0, 0
        pos += len(blk)
    d_src, d_dst = ctx.alloc(len(src)), ctx.alloc(cap * len(blocks))
    d_desc, d_st = ctx.alloc(ctypes_sizeof(descs)), ctx.alloc(24 * len(blocks))
    ctx.h2d(d_src, src)
    ctx.h2d(d_desc, bytes(descs))
    ctx.set_tuning(g)
    try:
        assert lz.lib().lz4b200_decode_blocks(ctx.handle, d_src, d_dst, len(blocks), d_desc, d_st) == 0
    finally:
        ctx.set_tuning(0)
    st = ctx.d2h(d_st, 24 * len(blocks))
    out = ctx.d2h(d_dst, cap * len(blocks))
    for i, exp in enumerate(expect):
        code, out_len = struct.unpack_from("<II", st, 24 * i)
        assert code == 0 and out_len == len(exp), (i, code, out_len, len(exp))
        o = descs[i].dst_off
        assert out[o:o + out_len] == exp, i
    for p in (d_src, d_dst, d_desc, d_st):
        ctx.free(p)



def _run_blocks_direct(ctx, blocks, cap, g, misalign=True, flags=0):
    """lz4b200_decode_blocks on raw blocks; returns [(code, out_len, bytes)]."""
    stride = (cap + 64 + 255) & ~255
    src = b"".join(blocks)
    descs = (lz.BlkDesc * len(blocks))()
    pos = 0
    for i, blk in enumerate(blocks):
        descs[i].src_off, descs[i].src_len = pos, len(blk)
        descs[i].dst_off, descs[i].dst_cap = i * stride + ((i % 16) if misalign else 0), cap
        descs[i].flags, descs[i].hist_avail = flags, 0
        pos += len(blk)
    d_src, d_dst = ctx.alloc(len(src)), ctx.alloc(stride * len(blocks))
    d_desc, d_st = ctx.alloc(ctypes_sizeof(descs)), ctx.alloc(24 * len(blocks))
    ctx.h2d(d_src, src)
    ctx.h2d(d_desc, bytes(descs))
    ctx.set_tuning(g)
    try:
        assert lz.lib().lz4b200_decode_blocks(ctx.handle, d_src, d_dst, len(blocks), d_desc, d_st) == 0
    finally:
        ctx.set_tuning(0)
    st = ctx.d2h(d_st, 24 * len(blocks))
    out = ctx.d2h(d_dst, stride * len(blocks))
    res = []
    for i in range(len(blocks)):
        code, out_len = struct.unpack_from("<II", st, 24 * i)
        o = descs[i].dst_off
        res.append((code, out_len, out[o:o + out_len]))
    for p in (d_src, d_dst, d_desc, d_st):
        ctx.free(p)
    return res


def _v3_shape_blocks():
    """Blocks aimed at the v3 kernel's machinery: several compressed windows per block, token-table
    cuts (dense 3-byte sequences), long literal runs (inside one window, across the window edge, larger
    than a window), long and overlapping matches, every literal / match length around 15 and 24."""
    import ctypes
    rng = np.random.default_rng(77)
    blocks = []
    # encoder-made blocks: text (one window), low-ratio text + noise (two windows), rle, short periods
    text = corpus.text_like(65536, seed=41)
    noisy = bytearray(text)
    for i in range(0, 65536, 3):
        noisy[i] = int(rng.integers(0, 256))
    for data in (text, bytes(noisy), corpus.rle_like(65536, seed=42), text[:100], text[:33000] + bytes(noisy[:32000]),
                 bytes(65536), (b"ab" * 40000)[:65536], corpus.random_bytes(20000, seed=3) + text[:45000]):
        blk = corpus.compress_block(data)
        if blk is not None and len(blk) < len(data):
            blocks.append(blk)
    # dense sequences: 0 literals + 4-byte matches -> 3 compressed bytes per sequence (token-table cut)
    seed = bytes(rng.integers(0, 256, 64, dtype=np.uint8))
    seqs = [(seed, 64, 4)]
    for i in range(15000):
        seqs.append((b"", int(rng.integers(1, 60)), 4))
    blocks.append(_raw_block(seqs, b"end"))
    # 1-literal sequences with near offsets (chains of in-batch dependencies)
    seqs = [(seed, 3, 9)]
    for i in range(9000):
        seqs.append((bytes([i & 255]), int(rng.integers(1, 12)), int(rng.integers(4, 7))))
    blocks.append(_raw_block(seqs, b""))
    # literal and match lengths around the nibble / LIT_EMIT / 32 thresholds, far and near offsets
    seqs = [(bytes(rng.integers(0, 256, 300, dtype=np.uint8)), 200, 40)]
    total = 340
    for ll in list(range(0, 40)) + [254, 255, 256, 269, 270, 271, 525, 1000]:
        for ml in (4, 18, 19, 20, 32, 33, 34, 273, 274, 275, 600):
            if total + ll + ml > 60000:
                continue
            off = int(rng.integers(1, min(total + ll, 65535) + 1))
            seqs.append((bytes(rng.integers(0, 256, ll, dtype=np.uint8)), off, ml))
            total += ll + ml
    blocks.append(_raw_block(seqs, b"0123456789abcdef"))
    # literal runs: 30000 (inside the first window), across the 33792-byte window edge, 40000 (> window: exact path)
    for first, run in ((100, 30000), (33000, 2000), (20, 40000)):
        pre = bytes(rng.integers(0, 4, first, dtype=np.uint8))
        lits = bytes(rng.integers(0, 256, run, dtype=np.uint8))
        seqs = [(pre, 1, 8)] + [(b"x", 2, 5)] * 50 + [(lits, 77, 12)] + [(b"yz", 3, 300)] * 10
        blocks.append(_raw_block(seqs, b"fin"))
    # one giant match per offset (RLE-like), and a match that ends the block without final literals
    for off in (1, 2, 3, 4, 7, 16, 31, 32, 33, 100, 5000):
        lead = bytes(rng.integers(0, 256, off + 5, dtype=np.uint8))
        blocks.append(_raw_block([(lead, off, 65536 - len(lead) - 2)], b"zz"))
    blocks.append(_raw_block([(b"abcdefgh", 8, 20)], b"")[:-1])   # no final token: block ends after a match
    return blocks


@pytest.mark.parametrize("g", [64, 41, 44, 48, 50, 60, 61, 8, -1])
def test_k1_v3_shapes_direct(ctx, oracle, g):
    """The shapes above through lz4b200_decode_blocks, compared with the pure-Python decoder (and with
    each other across kernel generations); destinations at every 16-byte phase."""
    blocks = _v3_shape_blocks()
    expect = [_py_decode(b) for b in blocks]
    assert all(len(e) <= 65536 for e in expect)
    for (code, out_len, out), exp, blk in zip(_run_blocks_direct(ctx, blocks, 65536, g), expect, blocks):
        assert code == 0 and out_len == len(exp), (g, len(blk), code, out_len, len(exp))
        assert out == exp, (g, len(blk))


@pytest.mark.parametrize("g", [60, 61])
def test_k1_v6_long_length_fields_stay_in_the_lane(ctx, g):
    """Length fields with extension bytes of 255 (literal runs >= 270, matches >= 274; Decompress_Sequence's length
    loops, lib/lz4ada.adb:741-747 / :773-777) are decoded by the v6 lane itself -- lz4b200_k1_fallbacks stays at 0 --
    at the block start, behind a match, with the long match directly behind the token and behind literals (the trip
    that stops in front of the offset); a field of more than 64 extension bytes goes to the exact routine and is
    still right."""
    rng = np.random.default_rng(99)
    noise = lambda n: bytes(rng.integers(0, 256, n, dtype=np.uint8))
    blocks = []
    for ll in (270, 271, 524, 525, 526, 779, 1000, 4000, 16000):
        blocks.append(_raw_block([(noise(ll), 5, 9), (b"ab", 2, 30)], b"end"))            # at the block start
        blocks.append(_raw_block([(b"qrstu", 5, 9), (noise(ll), 100, 20)], noise(ll)))     # behind a match; and final
    for ml in (273, 274, 275, 528, 529, 530, 783, 1039, 5000, 16000):
        for lead in (1, 3, 6, 7, 8, 14, 15, 16, 40):
            blocks.append(_raw_block([(noise(lead), lead, ml), (b"", 300 if lead + ml > 300 else 1, ml)], b"xy"))
    plain = blocks
    expect = [_py_decode(b) for b in plain]
    cap = max(len(e) for e in expect)
    for (code, out_len, out), exp in zip(_run_blocks_direct(ctx, plain, cap, g), expect):
        assert code == 0 and out == exp, (g, code, out_len, len(exp))
    assert ctx.k1_fallbacks() == (0, 0)
    # beyond what a lane follows: 80 extension bytes
    far = [_raw_block([(noise(15 + 255 * 80 + 7), 9, 12)], b"z"), _raw_block([(b"abc", 3, 19 + 255 * 80 + 3)], b"z")]
    expect = [_py_decode(b) for b in far]
    for (code, out_len, out), exp in zip(_run_blocks_direct(ctx, far, 24000, g), expect):
        assert code == 0 and out == exp, (g, code, out_len, len(exp))
    assert ctx.k1_fallbacks() == (2, 0)


def test_k1_v3_many_blocks_property(ctx):
    """4096 text blocks of 64 KiB (eight per CTA, every hash quad busy) through the batch path:
    output identical to the plain data and the device content checksums agree with the frames'."""
    data = corpus.text_like(1 << 20, seed=5) * 1
    frames = [corpus.build_frame(data[(i % 13) * 1000:] + data[:(i % 13) * 1000], 4, True, True) for i in range(64)]
    res = lz.batch_decompress(ctx, frames)
    for i, (exc, out, eof, msg) in enumerate(res):
        assert exc == "OK", (i, msg)
        assert out == data[(i % 13) * 1000:] + data[:(i % 13) * 1000], i


@pytest.mark.parametrize("g", [50, 60, 61, 41])
def test_k1_lane_refill_and_checksums_direct(ctx, oracle, g):
    """300 blocks from 1 byte to 64 KiB (text, RLE, noise) with block checksums through lz4b200_decode_blocks:
    v5 lanes finish at very different times and pull new blocks from the counter; every third block's checksum
    trailer is corrupted and must come back as LZ4B200_ST_BLOCK_CHECKSUM with the right computed value."""
    rng = np.random.default_rng(5)
    text = corpus.text_like(400000, seed=77)
    rle = corpus.rle_like(200000, seed=78)
    blocks, expect, bad = [], [], []
    for i in range(300):
        n = int(rng.choice([1, 7, 64, 300, 2000, 9000, 30000, 65536]))
        kind = i % 3
        base = text if kind == 0 else rle if kind == 1 else bytes(rng.integers(0, 256, 70000, dtype=np.uint8))
        o = int(rng.integers(0, len(base) - n))
        plain = base[o:o + n]
        blk = corpus.compress_block(plain)
        h = corpus.xxh32(blk)
        corrupt = i % 3 == 2 and i % 2 == 0
        blocks.append(blk + struct.pack("<I", h ^ (0x10 if corrupt else 0)))
        expect.append(plain)
        bad.append(corrupt)
    cap = 65536
    stride = cap + 256
    src = b"".join(blocks)
    descs = (lz.BlkDesc * len(blocks))()
    pos = 0
    for i, blk in enumerate(blocks):
        descs[i].src_off, descs[i].src_len = pos, len(blk) - 4
        descs[i].dst_off, descs[i].dst_cap = i * stride + (i % 5), cap
        descs[i].flags, descs[i].hist_avail = 2, 0   # LZ4B200_BLK_HAS_CHECKSUM
        pos += len(blk)
    d_src, d_dst = ctx.alloc(len(src)), ctx.alloc(stride * len(blocks))
    d_desc, d_st = ctx.alloc(ctypes_sizeof(descs)), ctx.alloc(24 * len(blocks))
    ctx.h2d(d_src, src)
    ctx.h2d(d_desc, bytes(descs))
    ctx.set_tuning(g)
    try:
        assert lz.lib().lz4b200_decode_blocks(ctx.handle, d_src, d_dst, len(blocks), d_desc, d_st) == 0
    finally:
        ctx.set_tuning(0)
    st = ctx.d2h(d_st, 24 * len(blocks))
    out = ctx.d2h(d_dst, stride * len(blocks))
    for i, (blk, exp, corrupt) in enumerate(zip(blocks, expect, bad)):
        code, out_len, _, _, computed, declared = struct.unpack_from("<IIIiII", st, 24 * i)
        assert computed == corpus.xxh32(blk[:-4]), i
        assert declared == struct.unpack("<I", blk[-4:])[0], i
        if corrupt:
            assert code == 1 and out_len == 0, (i, code)          # LZ4B200_ST_BLOCK_CHECKSUM
        else:
            assert code == 0 and out_len == len(exp), (i, code, out_len, len(exp))
            o = descs[i].dst_off
            assert out[o:o + out_len] == exp, i
    for p in (d_src, d_dst, d_desc, d_st):
        ctx.free(p)

def test_py_decoder_agrees_on_vector(oracle):
    """Sanity of the test-side helper against a golden vector (keeps _py_decode honest)."""
    frame = _read("z100.lz4")
    assert _py_decode(frame[11:11 + 11]) == _read("z100.bin")


def test_property_roundtrip_large(ctx):
    """Size-independent property at a larger size than the oracle is run on: encode -> GPU decode
    -> byte compare, plus the device content checksum agreeing with the encoder's (a checksum of
    checksums over 64 frames x 1 MiB, 64 KiB blocks, block + content checksums)."""
    c = corpus.build_corpus(64 << 20, 1 << 20, 4, kinds=("text", "rle", "random"), keep_plain=True)
    b = lz.Batch(ctx, c["src"], c["items"])
    d_src, d_dst = ctx.alloc(len(c["src"])), ctx.alloc(b.output_bytes)
    b.upload(d_src)
    b.run(d_src, d_dst)
    total = 0
    for r, plain in zip(b.results(), c["plain"]):
        assert r["exception"] == "OK", r
        assert r["out_len"] == len(plain)
        assert ctx.d2h(d_dst + r["dst_off"], r["out_len"]) == plain
        total += r["out_len"]
    t = b.traffic()
    assert t["decompressed_written"] == total == t["checksum_reread"]
    b.close()
    ctx.free(d_src)
    ctx.free(d_dst)


def test_pipelined_host_path(ctx, oracle):
    """lz4ada_batch_run_pipelined (the device stage of the e2e call, chunks on separate CUDA streams)
    and the one-shot lz4ada_batch_decompress: good, slow-path and corrupted streams mixed."""
    import ctypes
    c = corpus.build_corpus(24 << 20, 1 << 20, 4, kinds=("text", "rle", "random"), keep_plain=True)
    text = corpus.text_like(200000, seed=77)
    extra = [corpus.build_frame(text, 4, False, True, block_size=30000),              # short interior blocks
             corpus.build_frame(text, 4, True, True, independent=False),              # linked
             _read("corruptedcntchcksm.err"), _read("backrefoverflow.err"), _read("t300k.lz4")]
    streams = [bytes(c["src"][o:o + n]) for o, n in c["items"]] + extra
    plains = list(c["plain"]) + [text, text, None, None, _read("t300k.bin")]
    src = b"".join(streams)
    offs, pos = [], 0
    for st in streams:
        offs.append((pos, len(st)))
        pos += len(st)
    # (a) explicit pipelined run, 5 chunks
    b = lz.Batch(ctx, src, offs)
    need = b.output_bytes
    d_src, d_dst = ctx.alloc(len(src)), ctx.alloc(need)
    host_out = bytearray(need + 64)
    addr_out = (ctypes.c_uint8 * len(host_out)).from_buffer(host_out)
    assert lz.lib().lz4ada_batch_upload(b._h, None, None) == 0
    rc = lz.lib().lz4ada_batch_run_pipelined(b._h, b.src_addr, addr_out, d_src, d_dst, 5)
    assert rc == 0
    for st, plain, r in zip(streams, plains, b.results()):
        oexc, oout, oeof, omsg = oracle.decode_stream(st, chunk=0, out_cap=(len(plain) if plain else 1 << 16) + 64)
        assert (r["exception"], r["message"]) == (oexc, omsg)
        assert bytes(host_out[r["dst_off"]:r["dst_off"] + r["out_len"]]) == oout
        if plain is not None:
            assert oout == plain
    b.close()
    ctx.free(d_src)
    ctx.free(d_dst)
    # (b) the one-shot public call
    items = (lz.BatchItem * len(offs))()
    for k, (o, n) in enumerate(offs):
        items[k].src_off, items[k].src_len = o, n
    results = (lz.BatchResult * len(offs))()
    msgs = ctypes.create_string_buffer(256 * len(offs))
    out2 = bytearray(need + 64)
    addr2 = (ctypes.c_uint8 * len(out2)).from_buffer(out2)
    rc = lz.lib().lz4ada_batch_decompress(ctx.handle, src, len(src), addr2, need, len(offs), items,
                                          lz.RESERVATIONS["For_All"], results, msgs, 256)
    assert rc == 0
    for k, (st, plain) in enumerate(zip(streams, plains)):
        oexc, oout, oeof, omsg = oracle.decode_stream(st, chunk=0, out_cap=(len(plain) if plain else 1 << 16) + 64)
        assert lz.EXC_NAMES[results[k].exception] == oexc
        assert msgs.raw[256 * k:256 * (k + 1)].split(b"\0")[0].decode() == omsg
        assert bytes(out2[results[k].dst_off:results[k].dst_off + results[k].out_len]) == oout


@pytest.mark.parametrize("mode", [[], ["--update"]])
def test_cli_unlz4ada_b200(ctx, mode):
    """The stdin -> stdout tool (counterpart of tool_unlz4ada / tool_unlz4ada_simple; test_run.sh's check: rv = 0 and
    sha256(out) == sha256(.bin)) on every good vector -- through the batched device entry point (default) and through
    the Update loop (--update) -- and a non-zero exit with the library's exception text on an error vector."""
    import subprocess
    exe = os.path.join(ROOT, "tools", "unlz4ada_b200")
    if not os.path.exists(exe):
        pytest.skip("tools/unlz4ada_b200 not built")
    for stem in GOOD:
        p = subprocess.run([exe] + mode, input=_read(stem + ".lz4"), capture_output=True)
        assert p.returncode == 0, (stem, p.stderr)
        _check_output(stem, p.stdout)
    p = subprocess.run([exe] + mode, input=_read("corruptionoffset0.err"), capture_output=True)
    assert p.returncode == 1
    assert p.stderr.decode().strip() == "raised LZ4ADA.DATA_CORRUPTION : Corrupted Block: Offset = 0 detected."
    p = subprocess.run([exe] + mode, input=_read("t100k.lz4")[:5000], capture_output=True)
    assert p.returncode == 2
    # bytes decoded before an error are written first (corruptedcntchcksm: the whole content, then the checksum error)
    p = subprocess.run([exe] + mode, input=_read("corruptedcntchcksm.err"), capture_output=True)
    assert p.returncode == 1 and "content checksum" in p.stderr.decode() and len(p.stdout) > 0
    # -v: an I/O-inclusive figure on stderr
    p = subprocess.run([exe] + mode + ["-v"], input=_read("t1111k.lz4"), capture_output=True)
    assert p.returncode == 0 and "MB/s decompressed" in p.stderr.decode()
    _check_output("t1111k", p.stdout)


def test_pipelined_chains_repeatable(ctx):
    """The pipelined chain kernel (one parser warp feeding seven copier warps through a shared-memory
    ring) orders its batches with flags, not barriers: run linked frames and big solo blocks several
    times and demand the same exact bytes every time."""
    text = corpus.text_like(3 << 20, seed=123)
    rle = corpus.rle_like(1 << 20, seed=124)
    mix = text[:900000] + rle[:300000] + text[900000:1500000] + corpus.random_bytes(100000, seed=9) + text[1500000:2500000]
    streams, plains = [], []
    for i in range(12):
        p = mix[i * 1000:i * 1000 + 1500000 + i * 50000]
        streams.append(corpus.build_frame(p, 4 + (i % 4), i % 2 == 0, True, True, independent=False))   # linked
        plains.append(p)
    for i in range(6):
        p = text[i * 3000:i * 3000 + 2500000]
        streams.append(corpus.build_frame(p, 7, False, True))                                          # solo 4 MiB blocks
        plains.append(p)
        streams.append(corpus.build_legacy_frame(p))                                                    # legacy 8 MiB block
        plains.append(p)
    for rep in range(4):
        for plain, (exc, out, eof, msg) in zip(plains, lz.batch_decompress(ctx, streams)):
            assert exc == "OK", msg
            assert out == plain, rep


def _run_chains_direct(ctx, chains_of_blocks, cap_per_block, checksums=False, corrupt=()):
    """lz4b200_decode_linked on hand-made chains (lists of raw blocks): the blocks of a chain are one frame, output back
    to back.  Returns per chain [(code, out_len, computed, declared)] and the chain's output bytes."""
    src = bytearray(b"\0" * 16)
    descs, chains = [], []
    dst_pos = 0
    for blocks in chains_of_blocks:
        first = len(descs)
        for i, blk in enumerate(blocks):
            d = lz.BlkDesc()
            d.src_off, d.src_len = len(src), len(blk)
            d.dst_off, d.dst_cap = dst_pos, cap_per_block
            d.flags = 8 | (16 if i == 0 else 0) | (2 if checksums else 0)   # CHAINED, FIRST_OF_FRAME, HAS_CHECKSUM
            d.hist_avail = 0
            src += blk
            if checksums:
                h = corpus.xxh32(blk) ^ (1 if len(descs) in corrupt else 0)
                src += struct.pack("<I", h)
            src += b"\0" * (len(descs) % 5)   # every source alignment
            descs.append(d)
        c = lz.Chain()
        c.first_block, c.n_blocks = first, len(blocks)
        c.dst_off, c.dst_cap = dst_pos + 3 * len(chains), cap_per_block * len(blocks)   # every destination alignment
        chains.append(c)
        dst_pos += cap_per_block * len(blocks) + 256
    src += b"\0" * 64
    da = (lz.BlkDesc * len(descs))(*descs)
    ca = (lz.Chain * len(chains))(*chains)
    d_src, d_dst = ctx.alloc(len(src)), ctx.alloc(dst_pos + 256)
    d_desc, d_st, d_ch = ctx.alloc(ctypes_sizeof(da)), ctx.alloc(24 * len(descs)), ctx.alloc(ctypes_sizeof(ca))
    ctx.h2d(d_src, bytes(src))
    ctx.h2d(d_desc, bytes(da))
    ctx.h2d(d_ch, bytes(ca))
    assert lz.lib().lz4b200_memset(ctx.handle, d_st, 0xff, 24 * len(descs)) == 0
    assert lz.lib().lz4b200_decode_linked(ctx.handle, d_src, d_dst, len(chains), d_ch, d_desc, d_st) == 0
    ctx.sync()
    st = ctx.d2h(d_st, 24 * len(descs))
    out = ctx.d2h(d_dst, dst_pos + 256)
    res = []
    for c in chains:
        rows = [struct.unpack_from("<IIIiII", st, 24 * (c.first_block + i)) for i in range(c.n_blocks)]
        n = sum(r[1] for r in rows if r[0] == 0)
        res.append((rows, out[c.dst_off:c.dst_off + n]))
    for p_ in (d_src, d_dst, d_desc, d_st, d_ch):
        ctx.free(p_)
    return res


@pytest.mark.parametrize("sizes", [[5], [16, 1, 15, 17, 64], [65536] * 6 + [12345], [300000, 7, 65536, 200001], [1000] * 255])
def test_stream_adopt_list_direct(ctx, sizes):
    """lz4b200_stream_adopt_list (what Update's read-ahead hands the stream: up to 255 served blocks at once, one hash
    launch through K3's ring): the running content checksum equals XXH32 of the pieces in order (lengths that are not
    multiples of 16 exercise the carry between pieces), also when a second list follows; the per-block entry
    lz4b200_stream_adopt gives the same digest."""
    import ctypes
    rng = np.random.default_rng(len(sizes) * 7 + sizes[0])
    L = lz.lib()
    stride = 300032
    blob = bytearray(stride * len(sizes) + 64)
    offs, whole = [], bytearray()
    for i, n in enumerate(sizes):
        piece = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        off = i * stride + (i % 7)          # every alignment
        blob[off:off + n] = piece
        offs.append(off)
        whole += piece
    d = ctx.alloc(len(blob))
    ctx.h2d(d, bytes(blob))
    digests = []
    for mode in ("list", "list-twice", "single"):
        st = ctypes.c_void_p()
        assert L.lz4b200_stream_create(ctx.handle, 300000, ctypes.byref(st)) == 0
        arr_o = (ctypes.c_uint32 * len(sizes))(*offs)
        arr_n = (ctypes.c_uint32 * len(sizes))(*sizes)
        if mode == "single":
            for o, n in zip(offs, sizes):
                assert L.lz4b200_stream_adopt(st, ctypes.c_void_p(d + o), n, 1) == 0
        else:
            for _ in range(2 if mode == "list-twice" else 1):
                assert L.lz4b200_stream_adopt_list(st, ctypes.c_void_p(d), len(sizes), arr_o, arr_n, 1) == 0
        h = ctypes.c_uint32(0)
        assert L.lz4b200_stream_digest(st, ctypes.byref(h)) == 0
        digests.append(h.value)
        assert L.lz4b200_stream_destroy(st) == 0
    ctx.free(d)
    assert digests[0] == corpus.xxh32(bytes(whole))
    assert digests[1] == corpus.xxh32(bytes(whole) * 2)
    assert digests[2] == digests[0]


def test_chain_kernel_shapes_direct(ctx):
    """The chain kernel (K7 by default) straight through lz4b200_decode_linked: the hand-made shapes of the K1 tests as
    chains of one block (dense 3-byte sequences, literal runs of 30 000 / 40 000 bytes -- longer than the staged bytes --
    giant matches at every small offset, lengths around every nibble / extension threshold), multi-block chains whose
    matches reach back across block boundaries, every source and destination alignment, block checksums hashed by the
    kernel in front (one of them wrong: that block and the rest of its chain must not be reported as decoded)."""
    rng = np.random.default_rng(123)
    shapes = _v3_shape_blocks()
    ctx.chain_stats()
    res = _run_chains_direct(ctx, [[b] for b in shapes], 65536 + 64)
    for blk, (rows, out) in zip(shapes, res):
        exp = _py_decode(blk)
        assert rows[0][0] == 0 and rows[0][1] == len(exp), (len(blk), rows[0])
        assert out == exp, len(blk)
    if not os.environ.get("LZ4B200_CHAIN_KERNEL"):
        assert ctx.chain_stats() == (len(shapes), 0)   # all of them on K7's fast path, none through the exact routine
    # linked chains: one long text compressed block by block against the previous 64 KiB
    text = corpus.text_like(900000, seed=11) + bytes(70000) + corpus.random_bytes(5000, seed=2) + corpus.text_like(100000, seed=12)
    chains, plains = [], []
    for bs in (65536, 262144, 20000):
        blocks, pos = [], 0
        while pos < len(text):
            piece = text[pos:pos + bs]
            blocks.append(corpus.compress_block(piece, prefix=text[max(0, pos - 65536):pos]))
            pos += bs
        if all(b is not None for b in blocks):
            chains.append(blocks)
            plains.append(text)
    assert chains
    for checks in (False, True):
        bad = (3,) if checks else ()
        res = _run_chains_direct(ctx, chains, 262144 + 64, checksums=checks, corrupt=bad)
        for k, ((rows, out), plain) in enumerate(zip(res, plains)):
            if checks and k == 0:
                # block 3 of the first chain carries a wrong checksum (lib/lz4ada.adb:698-707): reported there, nothing after it
                assert [r[0] for r in rows[:3]] == [0, 0, 0] and rows[3][0] == 1 and rows[3][4] != rows[3][5], rows[:5]
                assert all(r[0] != 0 for r in rows[3:])
                assert out == plain[:len(out)] and len(out) == 3 * 65536
            else:
                assert all(r[0] == 0 for r in rows), (k, [r[0] for r in rows])
                assert out == plain, k
        if not os.environ.get("LZ4B200_CHAIN_KERNEL"):
            assert ctx.chain_stats()[1] == (1 if checks else 0)


@pytest.mark.parametrize("chain,solo", [("k7", "1"), ("k7", "0"), ("k6", "1"), ("k6", "0"), ("warp", "1"), ("pipe", "1")])
def test_chain_kernel_variants(chain, solo):
    """The chain kernels are chosen once per process from the environment: K6 (pointer-doubling parser + round-based
    copier on a shared-memory window), the K4 pipeline and the one-warp kernel, with and without big independent blocks
    placed as chains of one.  Each in a process of its own: good vectors, synthetic frames (linked, stored blocks
    inside linked frames, long literal runs, zero pages) and corrupted streams against the oracle."""
    import subprocess
    import sys
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import bo_lz4_ada_b200 as lz
import oracle_binding
import test_gpu_parity as T
from tools import corpus
oracle = oracle_binding.load()
ctx = lz.DeviceContext(0); ctx.make_default()
streams = [T._read(s + ".lz4") for s in T.GOOD]
for stem, (exc, out, eof, msg) in zip(T.GOOD, lz.batch_decompress(ctx, streams)):
    assert exc == "OK", (stem, msg)
    T._check_output(stem, out)
cases = T._synthetic_frames()
for (name, frame, plain), (exc, out, eof, msg) in zip(cases, lz.batch_decompress(ctx, [c[1] for c in cases])):
    assert exc == "OK" and out == plain, (name, msg)
text = corpus.text_like(3 << 20, seed=41)
big = [corpus.build_frame(text[i * 1000:i * 1000 + 2500000], 7, i %% 2 == 0, True) for i in range(4)]
big += [corpus.build_frame(text[:1500000] + bytes(300000) + corpus.random_bytes(200000, seed=3) + text[1500000:2000000], 6, True, True, independent=False)]
for k, (exc, out, eof, msg) in enumerate(lz.batch_decompress(ctx, big)):
    oexc, oout, oeof, omsg = oracle.decode_stream(big[k], chunk=0, out_cap=4 << 20)
    assert (exc, out) == (oexc, oout) and exc == "OK", (k, msg)
rng = np.random.default_rng(5)
mix = text[:200000] + bytes(5000) + corpus.random_bytes(3000, seed=1) + text[200000:450000]
bad = T._mutations(corpus.build_frame(mix, 4, True, True, True, independent=False), rng, 60) + T._mutations(big[1][:400000] + big[1][400000:], rng, 12)
for k, (data, (exc, out, eof, msg)) in enumerate(zip(bad, lz.batch_decompress(ctx, bad))):
    oexc, oout, oeof, omsg = oracle.decode_stream(data, chunk=0, out_cap=4 << 20)
    assert (exc, msg, out) == (oexc, omsg, oout), (k, exc, msg, oexc, omsg)
print("variant ok")
''' % (ROOT, ROOT)
    env = dict(os.environ, LZ4B200_CHAIN_KERNEL=chain, LZ4B200_SOLO=solo)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and "variant ok" in p.stdout, p.stderr[-2000:]


def test_batch_decompress_multi(oracle):
    """lz4ada_batch_decompress_multi: whole streams dealt to several device contexts in contiguous runs, every context
    on a host thread of its own, one source and one destination buffer (two and three contexts on the device this box
    has): outcomes, messages and bytes are the single-context call's."""
    import ctypes
    c = corpus.build_corpus(20 << 20, 1 << 20, 4, kinds=("text", "rle", "random"), keep_plain=True)
    text = corpus.text_like(200000, seed=78)
    extra = [corpus.build_frame(text, 4, False, True, block_size=30000), corpus.build_frame(text, 5, True, True, independent=False),
             _read("corruptedcntchcksm.err"), _read("cntblkszoverflow.err"), _read("t300k.lz4"), _read("z9m.lz4")]
    streams = [bytes(c["src"][o:o + n]) for o, n in c["items"]] + extra
    src = b"".join(streams)
    offs, pos = [], 0
    for st in streams:
        offs.append((pos, len(st)))
        pos += len(st)
    expect = [oracle.decode_stream(st, chunk=0, out_cap=10 << 20) for st in streams]
    cap = sum(len(e[1]) for e in expect) + (32 << 20)
    for n_ctx in (1, 2, 3):
        ctxs = [lz.DeviceContext(0) for _ in range(n_ctx)]
        handles = (ctypes.c_void_p * n_ctx)(*[cx.handle for cx in ctxs])
        items = (lz.BatchItem * len(offs))()
        for k, (o, n) in enumerate(offs):
            items[k].src_off, items[k].src_len = o, n
        results = (lz.BatchResult * len(offs))()
        msgs = ctypes.create_string_buffer(256 * len(offs))
        out = bytearray(cap)
        addr = (ctypes.c_uint8 * len(out)).from_buffer(out)
        rc = lz.lib().lz4ada_batch_decompress_multi(n_ctx, handles, src, len(src), addr, cap, len(offs), items,
                                                    lz.RESERVATIONS["For_All"], results, msgs, 256)
        assert rc == 0, (n_ctx, rc)
        for k, (oexc, oout, oeof, omsg) in enumerate(expect):
            assert lz.EXC_NAMES[results[k].exception] == oexc, (n_ctx, k)
            assert msgs.raw[256 * k:256 * (k + 1)].split(b"\0")[0].decode() == omsg, (n_ctx, k)
            assert bytes(out[results[k].dst_off:results[k].dst_off + results[k].out_len]) == oout, (n_ctx, k)
        # regions of different streams never overlap
        spans = sorted((results[k].dst_off, results[k].dst_off + results[k].out_len) for k in range(len(offs)) if results[k].out_len)
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])), n_ctx
        for cx in ctxs:
            cx.close()


@pytest.mark.parametrize("g", [-1, 8, 41, 50, 60, 61])
def test_guard_bands(ctx, g):
    """compute-sanitizer is closed on this pool, so out-of-bounds stores are hunted the plain way: the output buffer is
    painted, streams are placed by the CALLER with gaps between them and guard bands at both ends, and after the run
    every byte outside the bytes a stream produced must still carry the paint -- for each K1 generation, with stored
    blocks (K2), linked frames (K4), misaligned placements and streams that end in an error."""
    text = corpus.text_like(400000, seed=61)
    rle = corpus.rle_like(300000, seed=62)
    rnd = corpus.random_bytes(150000, seed=63)
    plains = [text, rle, rnd, text[:70001], text[1000:200000] + rle[:100000] + rnd[:70000], text[:65536 * 3]]
    streams = [corpus.build_frame(plains[0], 4, True, True), corpus.build_frame(plains[1], 4, False, True),
               corpus.build_frame(plains[2], 4, True, True), corpus.build_frame(plains[3], 5, True, True),
               corpus.build_frame(plains[4], 4, True, True, independent=False), corpus.build_frame(plains[5], 4, True, True)]
    streams.append(_read("backrefoverflow.err"))
    plains.append(None)
    src = b"".join(streams)
    guard = 4096
    items, pos, spos = [], guard + 3, 0
    for st, pl in zip(streams, plains):
        cap = (len(pl) if pl is not None else 65536) + 17
        items.append((spos, len(st), pos, cap))
        spos += len(st)
        pos += cap + 129 + (pos % 7)        # odd gaps: placements are not 16-byte aligned
    total = pos + guard
    ctx.set_tuning(g)
    try:
        b = lz.Batch(ctx, src, items)
        d_src, d_dst = ctx.alloc(len(src) + 64), ctx.alloc(total + 64)
        lz.lib().lz4b200_memset(ctx.handle, d_dst, 0xA5, total + 64)
        b.upload(d_src)
        b.run(d_src, d_dst)
        res = b.results()
        out = np.frombuffer(ctx.d2h(d_dst, total), dtype=np.uint8).copy()
        for (so, sl, do, dc), pl, r in zip(items, plains, res):
            if pl is not None:
                assert r["exception"] == "OK" and r["dst_off"] == do, r
                assert bytes(out[do:do + r["out_len"]]) == pl
            # erase what the stream legitimately produced (a stream that ends in an error may have written the bytes of
            # its failing block that precede the error: they are inside its region and are not handed out)
            out[do:do + (r["out_len"] if pl is not None else dc)] = 0xA5
        stray = np.flatnonzero(out != 0xA5)
        assert stray.size == 0, (g, stray[:8], len(stray))
        b.close()
        ctx.free(d_src)
        ctx.free(d_dst)
    finally:
        ctx.set_tuning(0)
