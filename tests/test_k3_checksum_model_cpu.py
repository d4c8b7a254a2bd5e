"""CPU models of the two things the content-checksum kernel (K3, csrc/kernels.cuh quad_xxh32_stream_t) changes
relative to a textbook XXH32 loop, checked against the oracle (XXHash32.Process, lib/lz4ada.adb:979-991):

1. the round in its carried-value forms (xxh_step): s = acc + x * P2, s' = rotl(s, 13) * P1 + x' * P2, and
   rotl(s, 13) * P1 = s * (P1 << 13) + (s >> 19) * P1 (mod 2^32);
2. the shared-memory ring schedule: groups of 16-byte granules requested with cp.async eight groups ahead, the
   batch of stripes of group g - 1 folded once group g has landed, the ring's first 32 bytes mirrored behind its end
   so that a batch reads linearly, and no granule requested that ends more than 19 bytes behind the span.

The kernel itself is tested on the GPU (tests/test_gpu_parity.py: test_k3_*); this file pins the arithmetic and
the index algebra where no GPU is needed."""
import numpy as np
import pytest

P1, P2 = 2654435761, 2246822519
M32 = 0xFFFFFFFF
GROUPS = 8


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def _finish(acc, n, tail):
    """Final, lib/lz4ada.adb:993-1017."""
    P3, P4, P5 = 3266489917, 668265263, 374761393
    h = (_rotl(acc[0], 1) + _rotl(acc[1], 7) + _rotl(acc[2], 12) + _rotl(acc[3], 18)) & M32 if n >= 16 else (acc[2] + P5) & M32
    h = (h + n) & M32
    while len(tail) >= 4:
        h = (_rotl((h + int.from_bytes(tail[:4], "little") * P3) & M32, 17) * P4) & M32
        tail = tail[4:]
    for b in tail:
        h = (_rotl((h + b * P5) & M32, 11) * P1) & M32
    h = ((h ^ (h >> 15)) * P2) & M32
    h = ((h ^ (h >> 13)) * P3) & M32
    return h ^ (h >> 16)


def _init():
    return [(P1 + P2) & M32, P2, 0, (0 - P1) & M32]


@pytest.mark.parametrize("form", ["shf_imad", "imadhi_imad"])
def test_carried_value_round_matches_oracle(oracle, form):
    rng = np.random.default_rng(3)
    for n in (16, 32, 48, 1000, 4096 + 7):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        acc = _init()
        for sub in range(4):
            xs = [int.from_bytes(data[s * 16 + sub * 4:s * 16 + sub * 4 + 4], "little") for s in range(n >> 4)]
            s = (acc[sub] + xs[0] * P2) & M32
            for x in xs[1:]:
                c = (x * P2) & M32
                if form == "shf_imad":
                    s = (_rotl(s, 13) * P1 + c) & M32
                else:
                    hi = (s * 8192) >> 32                      # mul.hi.u32 s, 2^13  ==  s >> 19
                    a = (s * ((P1 << 13) & M32) + c) & M32
                    s = (hi * P1 + a) & M32
            acc[sub] = (_rotl(s, 13) * P1) & M32
        assert _finish(acc, n, data[(n >> 4) << 4:]) == oracle.xxh32(data), (form, n)


def _ring_model(data, off, n, group):
    """quad_xxh32_stream_t, one quad, cp.async completion modelled as in the kernel (wait_group 6)."""
    ring_bytes, batch = GROUPS * group, group // 16
    ring = bytearray(ring_bytes + 48)
    mis = off & 15
    abase = off - mis
    nstripes = n >> 4
    need = ((mis + (nstripes << 4) + 4 + 15) & ~15) if nstripes else 0
    my_groups = (need + group - 1) // group
    sh = (mis & 3) * 8
    acc = _init()
    commits, done, furthest = [], 0, 0

    def issue(g):
        nonlocal furthest
        ops = []
        if g < my_groups:
            gb = g * group
            so = gb & (ring_bytes - 1)
            for sub in range(4):
                for c in range(group // 64):
                    if gb + c * 64 + sub * 16 < need:
                        ops.append((so + sub * 16 + c * 64, abase + gb + sub * 16 + c * 64))
                if so == 0 and sub < 2 and gb + sub * 16 < need:
                    ops.append((ring_bytes + sub * 16, abase + gb + sub * 16))
        for _, src in ops:
            furthest = max(furthest, src + 16)
        commits.append(ops)

    def wait(keep):
        nonlocal done
        while len(commits) - done > keep:
            for dst, src in commits[done]:
                ring[dst:dst + 16] = data[src:src + 16]
            done += 1

    for g in range(GROUPS):
        issue(g)
    n_batches = (nstripes + batch - 1) // batch
    for g in range(my_groups + 1):
        wait(GROUPS - 2)
        if g >= 1 and g - 1 < n_batches:
            s0 = (g - 1) * batch
            cnt = min(nstripes - s0, batch)
            for sub in range(4):
                b0 = (((g - 1) * group) & (ring_bytes - 1)) + (mis & ~3) + (sub << 2)
                for j in range(cnt):
                    w0 = int.from_bytes(ring[b0 + 16 * j:b0 + 16 * j + 4], "little")
                    w1 = int.from_bytes(ring[b0 + 16 * j + 4:b0 + 16 * j + 8], "little")
                    x = ((w0 | (w1 << 32)) >> sh) & M32
                    acc[sub] = (_rotl((acc[sub] + x * P2) & M32, 13) * P1) & M32
        if g >= 1:
            issue(g - 1 + GROUPS)
    return _finish(acc, n, data[off + (nstripes << 4):off + n]), furthest


@pytest.mark.parametrize("group", [256, 512, 1024, 2048])
def test_ring_schedule_model_matches_oracle(oracle, group):
    rng = np.random.default_rng(group)
    data = rng.integers(0, 256, 60000, dtype=np.uint8).tobytes()
    ring = GROUPS * group
    for off in (0, 1, 3, 4, 13, 15, 16, 4097):
        for n in (0, 15, 16, 17, 255, group, ring - 1, ring, ring + 16, ring + 33, 2 * ring + group + 5):
            got, furthest = _ring_model(data, off, n, group)
            assert got == oracle.xxh32(data[off:off + n]), (group, off, n)
            assert furthest <= off + n + 19 or n < 16, (group, off, n, furthest)   # the documented read bound
