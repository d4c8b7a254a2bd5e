"""Synthetic corpora and an LZ4 *frame writer* for the test and bench harness.

Harness-side only: nothing here is on the decode path.  Blocks are compressed with liblz4.so.1
through ctypes when it is loadable (LZ4_compress_default / LZ4_compress_fast_continue), otherwise
with the small greedy encoder in tools/lz4_greedy_enc.c.  Frames are assembled here so that every
frame feature the reference parses can be produced: block-max code, block checksums, content
checksum, content size, dictionary-id field, linked or independent blocks, stored blocks, legacy
frames, skippable frames.

Corpus classes (deterministic, numpy.random.default_rng(seed)):
  text_like   Zipf(1.15) over a 20 000-word vocabulary of lower-case words of length 2..9
              -> ratio ~2.2 at 64 KiB blocks, ~10 output bytes per sequence
  rle_like    zero runs and short-period (1,2,3,4,7,16) runs -> ratio ~250
  random      uniform bytes -> every block stored
"""
import ctypes
import os
import struct
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAGIC_MODERN = 0x184D2204
MAGIC_LEGACY = 0x184C2102
BLOCK_MAX = {4: 64 << 10, 5: 256 << 10, 6: 1 << 20, 7: 4 << 20}

# ----------------------------------------------------------------------------- XXH32 (harness)
try:
    import xxhash as _xxhash

    def xxh32(data):
        return _xxhash.xxh32(data, seed=0).intdigest()
except ImportError:   # pragma: no cover - the image has python-xxhash
    def xxh32(data):
        import bo_lz4_ada_b200 as pkg
        return pkg.XXHash32.Hash(bytes(data))

# ----------------------------------------------------------------------------- block encoders
_lz4 = None
_greedy = None
_tls = threading.local()


def _load_liblz4():
    global _lz4
    if _lz4 is None:
        try:
            L = ctypes.CDLL("liblz4.so.1")
            L.LZ4_compressBound.restype = ctypes.c_int
            L.LZ4_compressBound.argtypes = [ctypes.c_int]
            L.LZ4_compress_default.restype = ctypes.c_int
            L.LZ4_compress_default.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
            L.LZ4_createStream.restype = ctypes.c_void_p
            L.LZ4_freeStream.argtypes = [ctypes.c_void_p]
            L.LZ4_compress_fast_continue.restype = ctypes.c_int
            L.LZ4_compress_fast_continue.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                     ctypes.c_int, ctypes.c_int]
            _lz4 = L
        except OSError:
            _lz4 = False
    return _lz4


def _load_greedy():
    global _greedy
    if _greedy is None:
        so = os.path.join(_HERE, "liblz4greedy.so")
        src = os.path.join(_HERE, "lz4_greedy_enc.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
        G = ctypes.CDLL(so)
        G.lz4g_bound.restype = ctypes.c_int
        G.lz4g_bound.argtypes = [ctypes.c_int]
        G.lz4g_compress.restype = ctypes.c_int
        G.lz4g_compress.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _greedy = G
    return _greedy


def encoder_name(prefer="liblz4"):
    return "liblz4-1.9.4(ctypes)" if prefer == "liblz4" and _load_liblz4() else "greedy(tools/lz4_greedy_enc.c)"


def compress_block(data, prefix=b"", encoder="liblz4"):
    """One LZ4 block.  prefix = up to 64 KiB of preceding data the block may reference (linked)."""
    data = bytes(data)
    n = len(data)
    L = _load_liblz4() if encoder == "liblz4" else False
    if L:
        cap = L.LZ4_compressBound(n)
        dst = ctypes.create_string_buffer(cap)
        if prefix:
            # prefix mode: history must sit directly in front of the block in memory
            buf = ctypes.create_string_buffer(bytes(prefix) + data, len(prefix) + n)
            base = ctypes.addressof(buf)
            st = L.LZ4_createStream()
            tmp = ctypes.create_string_buffer(L.LZ4_compressBound(len(prefix)))
            L.LZ4_compress_fast_continue(st, base, tmp, len(prefix), len(tmp), 1)
            k = L.LZ4_compress_fast_continue(st, base + len(prefix), dst, n, cap, 1)
            L.LZ4_freeStream(st)
        else:
            k = L.LZ4_compress_default(data, dst, n, cap)
        assert k > 0
        return dst.raw[:k]
    G = _load_greedy()
    cap = G.lz4g_bound(n)
    dst = ctypes.create_string_buffer(cap)
    buf = bytes(prefix) + data
    k = G.lz4g_compress(buf, len(prefix), n, dst, cap)
    assert k > 0
    return dst.raw[:k]


# ----------------------------------------------------------------------------- frame writer
def frame_header(bmax_code=4, block_checksum=False, content_checksum=True, content_size=None, independent=True,
                 dict_id=None):
    flg = (1 << 6) | (0x20 if independent else 0) | (0x10 if block_checksum else 0) | \
          (0x08 if content_size is not None else 0) | (0x04 if content_checksum else 0) | \
          (0x01 if dict_id is not None else 0)
    bd = bmax_code << 4
    desc = bytes([flg, bd])
    if content_size is not None:
        desc += struct.pack("<Q", content_size)
    if dict_id is not None:
        desc += struct.pack("<I", dict_id)
    hc = (xxh32(desc) >> 8) & 0xFF
    return struct.pack("<I", MAGIC_MODERN) + desc + bytes([hc])


def build_frame(data, bmax_code=4, block_checksum=False, content_checksum=True, content_size=False,
                independent=True, encoder="liblz4", block_size=None, dict_id=None, force_stored=False):
    """A complete modern frame.  block_size (<= block max) lets tests make short interior blocks."""
    data = bytes(data)
    bmax = BLOCK_MAX[bmax_code]
    bs = block_size or bmax
    out = [frame_header(bmax_code, block_checksum, content_checksum, len(data) if content_size else None,
                        independent, dict_id)]
    for off in range(0, len(data), bs):
        raw = data[off:off + bs]
        prefix = b"" if independent else data[max(0, off - 65536):off]
        comp = None if force_stored else compress_block(raw, prefix, encoder)
        if comp is None or len(comp) >= len(raw):
            payload, word = raw, len(raw) | 0x80000000
        else:
            payload, word = comp, len(comp)
        out.append(struct.pack("<I", word))
        out.append(payload)
        if block_checksum:
            out.append(struct.pack("<I", xxh32(payload)))
    out.append(struct.pack("<I", 0))
    if content_checksum:
        out.append(struct.pack("<I", xxh32(data)))
    return b"".join(out)


def build_legacy_frame(data, encoder="liblz4", block_size=8 << 20):
    data = bytes(data)
    out = [struct.pack("<I", MAGIC_LEGACY)]
    for off in range(0, len(data), block_size):
        comp = compress_block(data[off:off + block_size], b"", encoder)
        out.append(struct.pack("<I", len(comp)))
        out.append(comp)
    return b"".join(out)


def skippable_frame(payload=b"", nibble=0):
    return struct.pack("<II", 0x184D2A50 + nibble, len(payload)) + bytes(payload)


# ----------------------------------------------------------------------------- data classes
_VOCAB_CACHE = {}


def _vocab(seed):
    if seed not in _VOCAB_CACHE:
        rng = np.random.default_rng(seed)
        V = 20000
        lens = rng.integers(2, 10, V)
        M = np.zeros((V, 10), dtype=np.uint8)
        letters = rng.integers(97, 123, (V, 10), dtype=np.uint8)
        cols = np.arange(10)[None, :]
        M[:] = np.where(cols < lens[:, None], letters, 0)
        M[np.arange(V), lens] = 32   # trailing space
        p = 1.0 / np.arange(1, V + 1) ** 1.15
        cdf = np.cumsum(p / p.sum())
        _VOCAB_CACHE[seed] = (M, lens + 1, cdf)
    return _VOCAB_CACHE[seed]


def text_like(nbytes, seed=1234, vocab_seed=1234):
    """Zipf(1.15) text over a fixed 20 000-word vocabulary; exactly nbytes long."""
    M, wlen, cdf = _vocab(vocab_seed)
    rng = np.random.default_rng(seed)
    out = np.empty(nbytes, dtype=np.uint8)
    have = 0
    while have < nbytes:
        n = int((nbytes - have) / 6.0) + 64
        idx = np.searchsorted(cdf, rng.random(n))
        idx = np.minimum(idx, len(wlen) - 1)
        rows = M[idx]
        mask = np.arange(10)[None, :] < wlen[idx][:, None]
        chunk = rows[mask]
        k = min(len(chunk), nbytes - have)
        out[have:have + k] = chunk[:k]
        have += k
    return out.tobytes()


def rle_like(nbytes, seed=1):
    """Runs of short-period patterns (periods 1,2,3,4,7,16) of random lengths 1..64 KiB."""
    rng = np.random.default_rng(seed)
    out = np.empty(nbytes, dtype=np.uint8)
    pos = 0
    periods = [1, 1, 1, 2, 3, 4, 7, 16]
    while pos < nbytes:
        run = int(rng.integers(1024, 65536))
        run = min(run, nbytes - pos)
        p = periods[int(rng.integers(0, len(periods)))]
        pat = rng.integers(0, 256, p, dtype=np.uint8) if rng.random() < 0.7 else np.zeros(p, dtype=np.uint8)
        reps = (run + p - 1) // p
        out[pos:pos + run] = np.tile(pat, reps)[:run]
        pos += run
    return out.tobytes()


def random_bytes(nbytes, seed=2):
    return np.random.default_rng(seed).integers(0, 256, nbytes, dtype=np.uint8).tobytes()


GENERATORS = {"text": text_like, "rle": rle_like, "random": random_bytes}


# ----------------------------------------------------------------------------- whole corpora
def build_corpus(total_bytes, frame_bytes, bmax_code, kinds=("text",), block_checksum=True, content_checksum=True,
                 independent=True, seed=1234, workers=None, encoder="liblz4", keep_plain=False):
    """Many single-frame streams of `frame_bytes` each, cycling through `kinds`.

    Returns dict(src=bytearray of all frames back to back, items=[(off, len)], plain_bytes=total,
                 digests=[xxh32 of each frame's plain data], plain=[bytes] if keep_plain, kinds=[...]).
    """
    n_frames = max(1, total_bytes // frame_bytes)
    workers = workers or min(32, os.cpu_count() or 8)

    def one(i):
        kind = kinds[i % len(kinds)]
        data = GENERATORS[kind](frame_bytes, seed=seed + i)
        frame = build_frame(data, bmax_code, block_checksum, content_checksum, False, independent, encoder)
        return frame, xxh32(data), (data if keep_plain else None), kind

    with ThreadPoolExecutor(workers) as ex:
        results = list(ex.map(one, range(n_frames)))
    total = sum(len(r[0]) for r in results)
    src = bytearray(total)
    items, pos = [], 0
    for r in results:
        src[pos:pos + len(r[0])] = r[0]
        items.append((pos, len(r[0])))
        pos += len(r[0])
    return {"src": src, "items": items, "plain_bytes": n_frames * frame_bytes, "digests": [r[1] for r in results],
            "plain": [r[2] for r in results] if keep_plain else None, "kinds": [r[3] for r in results],
            "encoder": encoder_name(encoder)}
