"""CPU model of the K1 v6 lane (bo_lz4_ada_b200/csrc/kernels_v6.cuh): one PIECE of a sequence per trip.

The model follows the kernel's parse side statement by statement -- the 12-byte view at the cursor, the state
(rem_l, rem_m, need_off, dist, mln), a token's first piece carrying <= 7 literal bytes and a continuation <= 8, a match
piece of min(16, distance) bytes whose distance doubles while it is below 16 (pattern replication,
lib/lz4ada.adb:893-903), length fields with several extension bytes followed out of the in ring (<= 64 bytes of
them, otherwise the exact routine), a long match length only taken by a trip that starts at the offset -- and executes
the pieces in order on a byte array.  Checked against the plain data and a straight LZ4 decode: what the kernel's
descriptors say must be the block (Decompress_Sequence / Output_With_History, lib/lz4ada.adb:737-904).

The kernel itself is compared with the oracle on the GPU (tests/test_gpu_parity.py, every K1 tuning); this file pins
the trip logic, in particular the corners that once sent whole blocks to the exact routine."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import corpus  # noqa: E402

LIT_PIECE, ML_PIECE, VIEW = 8, 16, 12


class NeedsExactRoutine(Exception):
    pass


def v6_lane(block, cap):
    """-> (output bytes, number of trips that parsed something).  Raises NeedsExactRoutine where the kernel sets `bad`."""
    n = len(block)
    pad = bytes(block) + bytes(80)           # the in ring: bytes beyond the block are there, never used
    out = bytearray()
    a = 0
    rem_l = rem_m = need_off = dist = mln = 0
    ended = False
    trips = 0
    while not ended:
        trips += 1
        assert trips < 4 * n + cap + 64      # (a trip parses a token, or moves at least one byte)
        v = pad[a:a + VIEW]
        fresh = (rem_l | rem_m | need_off) == 0
        long_l = fresh and a + 1 < n and v[0] >= 0xf0 and v[1] == 255
        long_m = need_off != 0 and rem_l == 0 and mln == 15 and a + 2 < n and v[2] == 255
        x_sum = x_cnt = 0
        x_over = False
        if long_l or long_m:
            pos = a + (2 if long_l else 3)
            while True:
                if pos >= n or pos - a > 64:
                    x_over = True
                    break
                b = pad[pos]
                x_sum += b
                x_cnt += 1
                pos += 1
                if b != 255:
                    break
        end0 = fresh and a >= n
        tok = fresh and not end0
        tk, e1 = v[0], v[1]
        ext_l = tok and tk >= 0xf0
        o = ((2 + x_cnt) if ext_l else 1) if tok else 0
        if tok:
            rem_l = (tk >> 4) + ((e1 + x_sum) if ext_l else 0)
            mln = tk & 15
            need_off = 1
        if (ext_l and a + 1 >= n) or (long_l and x_over):
            raise NeedsExactRoutine("literal length field")
        if tok and rem_l > n - a - o:
            raise NeedsExactRoutine("literals beyond the block")
        lim = 0 if long_l else (7 if tok else LIT_PIECE)
        nl = min(rem_l, lim)
        lits = pad[a + o:a + o + nl]          # (the kernel: funnel shifts of the view -- o + nl <= 12 whenever nl > 0)
        assert nl == 0 or o + nl <= VIEW
        rem_l -= nl
        o += nl
        do_off = need_off != 0 and rem_l == 0 and not end0
        ao = a + o
        fin_lit = do_off and ao >= n
        assert not do_off or o + 3 <= VIEW     # offset and first extension byte inside the view
        off = pad[ao] | (pad[ao + 1] << 8)
        e2 = pad[ao + 2]
        ext_m = mln == 15
        defer = ext_m and e2 == 255 and not long_m and ao + 2 < n
        has_m = do_off and not fin_lit and not defer
        if has_m:
            rem_m = mln + 4 + ((e2 + x_sum) if ext_m else 0)
            dist = off
            o += (3 + x_cnt) if ext_m else 2
            need_off = 0
        if fin_lit and mln != 0:
            raise NeedsExactRoutine("match nibble on the final sequence")
        if has_m and (ao + 2 > n or off == 0):
            raise NeedsExactRoutine("offset")
        if has_m and off > len(out) + nl:
            raise NeedsExactRoutine("match before the block")
        if has_m and ext_m and (ao + 2 >= n or (long_m and x_over)):
            raise NeedsExactRoutine("match length field")
        if end0 or fin_lit:
            ended = True
            need_off = 0
        a += o
        cnt = min(rem_m, ML_PIECE, dist) if rem_m else 0
        rem_m -= cnt
        if nl + cnt > cap - len(out):
            raise NeedsExactRoutine("capacity")
        # ---- the copy side, K trips later: literals out of the descriptor, then the match piece ----
        out += lits
        if cnt:
            src = len(out) - dist
            out += out[src:src + cnt]          # cnt <= dist: never overlaps
        if dist < ML_PIECE and cnt == dist:
            dist <<= 1                         # the pattern has doubled
    return bytes(out), trips


def _py_decode(block):
    out, i = bytearray(), 0
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
        for _ in range(ml + 4):
            out.append(out[-off])
    return bytes(out)


def _raw_block(seqs, last_literals=b""):
    out = bytearray()

    def length(v):
        while v >= 255:
            out.append(255)
            v -= 255
        out.append(v)

    for lit, off, ml in seqs:
        ll, mm = len(lit), ml - 4
        out.append((min(ll, 15) << 4) | min(mm, 15))
        if ll >= 15:
            length(ll - 15)
        out += lit
        out += bytes([off & 255, off >> 8])
        if mm >= 15:
            length(mm - 15)
    ll = len(last_literals)
    out.append(min(ll, 15) << 4)
    if ll >= 15:
        length(ll - 15)
    out += last_literals
    return bytes(out)


@pytest.mark.parametrize("kind", ["text", "rle", "noise+text", "zeros"])
def test_encoder_made_blocks(kind):
    data = {"text": corpus.text_like(65536, seed=21), "rle": corpus.rle_like(65536, seed=22),
            "noise+text": corpus.random_bytes(3000, seed=5) + corpus.text_like(60000, seed=23), "zeros": bytes(65536)}[kind]
    blk = corpus.compress_block(data)
    try:
        out, trips = v6_lane(blk, len(data))
    except NeedsExactRoutine as why:
        # runs of 16 KiB and more (65 extension bytes and up) are beyond what a lane follows: zero pages and RLE data go
        # to the exact routine -- which is also why the batch scheduler does not count such blocks for v6
        assert kind in ("zeros", "rle") and "length field" in str(why)
        return
    assert out == data
    if kind == "text":
        assert trips < 1.3 * 6400            # about one trip per sequence (7 133 trips for ~6 300 sequences)


def test_length_fields_and_periods():
    rng = np.random.default_rng(4)
    noise = lambda k: bytes(rng.integers(0, 256, k, dtype=np.uint8))
    blocks = []
    for ll in (0, 1, 6, 7, 8, 9, 14, 15, 16, 23, 269, 270, 271, 524, 525, 526, 1000, 16000):
        for ml in (4, 18, 19, 20, 34, 273, 274, 275, 528, 529, 530, 5000):
            blocks.append(_raw_block([(noise(max(ll, 1)), max(ll, 1), 9), (noise(ll), 7, ml), (b"", 3, ml)], b"tail"))
    for off in list(range(1, 40)) + [255, 256, 257, 1000]:
        lead = noise(off + 3)
        blocks.append(_raw_block([(lead, off, 4), (b"", off, 700), (b"q", off, 15), (b"", 1, 16), (b"", 2, 17)], b""))
    blocks.append(_raw_block([(b"abcdefgh", 8, 20)], b"")[:-1])          # the block ends behind a match
    blocks.append(_raw_block([], noise(5000)))                            # literals only
    for blk in blocks:
        exp = _py_decode(blk)
        out, _ = v6_lane(blk, len(exp))
        assert out == exp
        with pytest.raises(NeedsExactRoutine):                            # one byte of room too few: the exact routine reports it
            v6_lane(blk, len(exp) - 1)


def test_what_goes_to_the_exact_routine():
    noise = bytes(range(256)) * 90
    too_long_l = _raw_block([(noise[:15 + 255 * 70 + 3], 9, 12)], b"z")   # 71 extension bytes
    too_long_m = _raw_block([(b"abc", 3, 19 + 255 * 70)], b"z")
    zero_off = _raw_block([(b"abc", 3, 8)], b"z").replace(b"\x03\x00", b"\x00\x00", 1)
    far_off = _raw_block([(b"abc", 4, 8)], b"z")
    truncated = _raw_block([(b"abcdef", 3, 8)], b"zzzz")[:-3]
    for blk in (too_long_l, too_long_m, zero_off, far_off, truncated):
        with pytest.raises(NeedsExactRoutine):
            v6_lane(blk, 1 << 20)
