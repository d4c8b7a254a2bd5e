"""CPU model of the K7 chain kernel's two ideas (bo_lz4_ada_b200/csrc/kernels_k7.cuh), checked against the plain data:

  * the speculative parse: every 32 / 64-byte segment of the compressed bytes is walked from its first byte as if a
    token started there; the true token chain is then threaded through the segments (neighbour rule, then the path
    from segment 0 by pointer jumping over "the segment my exit lands in") -- the fixed point must be exactly the
    block's token chain, also behind literal runs that jump over many segments;
  * the copy without a serial walk: literal bytes are final, a match byte is a pointer `offset` back, sources in front
    of the 16 KiB window are fetched; pointer jumping (a byte takes its target's byte when that is final, otherwise it
    adds the target's distance to its own) must end with the block's plain bytes, overlapping matches included.

The kernel itself is compared with the oracle on the GPU (tests/test_gpu_parity.py); this file pins the algorithm."""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import corpus  # noqa: E402


def _next_token(blk, x):
    """Decompress_Sequence's length arithmetic (lib/lz4ada.adb:737-777) -> (lit, ml, lit_pos, next) or None."""
    n = len(blk)
    t = blk[x]
    lit, ml, p = t >> 4, t & 15, x + 1
    if lit == 15:
        while True:
            if p >= n:
                return None
            e = blk[p]
            p += 1
            lit += e
            if e != 255:
                break
    lp, q = p, p + lit
    if q > n:
        return None
    if q == n:
        return (lit, 0, lp, n) if ml == 0 else None
    if q + 2 > n:
        return None
    p = q + 2
    if ml == 15:
        while True:
            if p >= n:
                return None
            e = blk[p]
            p += 1
            ml += e
            if e != 255:
                break
    return lit, ml + 4, lp, p


def _true_chain(blk):
    out, x = [], 0
    while x < len(blk):
        out.append(x)
        x = _next_token(blk, x)[3]
    return out


def _parse_step(blk, ip, seg, nseg):
    """One step of the kernel's parse: -> (token positions of the true chain inside the step, exit, iterations)."""
    n = len(blk)

    def walk(x0, t, old=None):
        seg_end = ip + (t + 1) * seg
        path, x = [], x0
        while True:
            if x >= n:
                return path, "END"
            if x >= seg_end:
                return path, x
            if old is not None and x in old[0]:
                return path + [p for p in old[1] if p >= x], old[2]
            tk = _next_token(blk, x)
            if tk is None:
                return path, "ERR"
            path.append(x)
            x = tk[3]

    state = []   # per segment: [set(path), path, exit, entry]
    for t in range(nseg):
        xt = ip + t * seg
        if xt >= n:
            state.append(None)
            continue
        p, e = walk(xt, t)
        state.append([set(p), p, e, xt])

    def seg_of(e):
        if not isinstance(e, int):
            return None
        k = (e - ip) // seg
        return k if k < nseg and state[k] is not None else None

    iterations, phase = 0, "neighbours"
    while True:
        iterations += 1
        exits = [s[2] if s else None for s in state]
        incoming, reached = [None] * nseg, [False] * nseg
        if phase == "path":
            t = 0
            while t is not None:
                reached[t] = True
                k = seg_of(exits[t])
                if k is not None:
                    incoming[k] = exits[t]
                t = k
        fresh = []
        for t in range(1, nseg):
            s = state[t]
            if s is None:
                continue
            xt = ip + t * seg
            if reached[t]:
                inc = incoming[t]
            else:
                inc = exits[t - 1] if isinstance(exits[t - 1], int) and xt <= exits[t - 1] < xt + seg else None
            if inc is not None and inc != s[3]:
                p, e = walk(inc, t, (s[0], s[1], s[2]))
                fresh.append((t, [set(p), p, e, inc]))
        for t, st in fresh:
            state[t] = st
        if not fresh:
            if phase == "neighbours":
                phase = "path"
            else:
                break
        assert iterations < 4 * nseg
    toks, t, last = [], 0, None
    while t is not None:
        toks += [p for p in state[t][1] if p >= state[t][3]]
        last = state[t][2]
        t = seg_of(last)
    return toks, last, iterations


def _mixed(nbytes, seed):
    """Text with incompressible stretches: literal runs that jump over many segments."""
    rng = random.Random(seed)
    text = corpus.text_like(nbytes, seed=seed)
    out, i = bytearray(), 0
    while len(out) < nbytes:
        k = rng.randrange(100, 3000)
        out += text[i:i + k]
        i += k
        out += corpus.random_bytes(rng.randrange(20, 900), seed=seed + i)
    return bytes(out[:nbytes])


@pytest.mark.parametrize("seg", [32, 64])
@pytest.mark.parametrize("kind", ["text", "mixed", "rle"])
def test_speculative_parse_reaches_the_true_chain(kind, seg):
    data = {"text": corpus.text_like(200000, seed=3), "mixed": _mixed(200000, 4), "rle": corpus.rle_like(200000, seed=5)}[kind]
    blk = corpus.compress_block(data)
    truth = _true_chain(blk)
    nseg = 64
    ip, steps, worst = 0, 0, 0
    got = []
    while ip < len(blk) and steps < 400:
        toks, last, iters = _parse_step(blk, ip, seg, nseg)
        got += toks
        worst = max(worst, iters)
        steps += 1
        if last in ("END", "ERR"):
            assert last == "END"
            break
        assert last > ip
        ip = last
    assert got == truth[:len(got)]
    assert steps == 400 or got == truth
    assert worst <= nseg   # (a dozen on text; the bound is what the kernel relies on for termination)


def _resolve_window(ptr, win):
    """Pointer jumping as kernels_k7.cuh does it: ptr[i] = 0 final, else distance to the byte it repeats."""
    rounds = 0
    while any(ptr):
        rounds += 1
        nptr = list(ptr)
        for i, d in enumerate(ptr):
            if d:
                j = i - d
                if ptr[j] == 0:
                    win[i] = win[j]
                    nptr[i] = 0
                else:
                    nptr[i] = d + ptr[j]
        ptr[:] = nptr
        assert rounds < 64
    return rounds


@pytest.mark.parametrize("kind", ["text", "rle", "mixed", "periods"])
def test_pointer_jumping_rebuilds_the_block(kind):
    if kind == "periods":
        data = b"".join(bytes([65 + k]) * (k + 1) + (b"ab" * 40)[:17 + k] + b"xyz" * k for k in range(120)) * 6
    else:
        data = {"text": corpus.text_like(70000, seed=13), "rle": corpus.rle_like(70000, seed=14), "mixed": _mixed(70000, 15)}[kind]
    blk = corpus.compress_block(data)
    nwin = 16384
    out = bytearray()
    win, ptr, w0 = bytearray(nwin), [0] * nwin, 0
    deepest = 0

    def finish(upto):
        nonlocal deepest
        deepest = max(deepest, _resolve_window(ptr, win))
        out.extend(win[len(out) - w0:upto - w0])

    pos = 0
    for x in _true_chain(blk):
        lit, ml, lp, _ = _next_token(blk, x)
        off = blk[lp + lit] | (blk[lp + lit + 1] << 8) if ml else 0
        for k in range(lit + ml):
            if pos - w0 == nwin:
                finish(pos)
                w0 += nwin
                ptr[:] = [0] * nwin
            el = pos - w0
            if k < lit:
                win[el] = blk[lp + k]           # a literal byte: final
            elif pos - off < w0:
                win[el] = out[pos - off]        # source in front of the window: global memory, final
            else:
                ptr[el] = off                   # a match byte: the byte `off` in front of it
            pos += 1
    finish(pos)
    assert bytes(out) == data
    assert deepest <= 15   # log2(window) + 1
