// kernels_v4.cuh -- K1 fourth generation: one warp per block, everything warp-synchronous.
//
// What the earlier generations measured:
//   v2  G blocks per warp, output assembled in global memory: 35 warp instructions per sequence and
//       7x the algorithmic DRAM traffic (16 000 blocks open at once, byte-granular stores to L2).
//   v3  a CTA per block with the whole 64 KiB output window in shared memory: only two blocks open per
//       SM, and the in-order hand-over between batches costs ~4 000 cycles per 32 sequences because a
//       lone warp issues a dependent instruction every ~5 cycles -- 5x slower than v2.
// The lesson: one block is a latency-bound stream (a batch cannot start before the previous one ends),
// so the SM needs 16+ independent streams, each cheap in instructions and each with its own *recent*
// output on chip.  v4 gives every warp its own block and ~10 KiB of shared memory:
//
//   cw      the next 4 224 compressed bytes (32 segments of 132), loaded with cp.async
//   parse   lane j walks segment j from a guessed token start; "my entry = left neighbour's exit" is
//           iterated to the fixed point with shuffles (exact; two or three walks on real data).  All
//           loops are warp-uniform with predicated bodies: no lane ever runs ahead of the others.
//   tokpos  token positions of the window, compact, in stream order (the transpose from "lane per
//           segment" to "lane per sequence")
//   ring    the last 4 KiB of output.  A batch of 32 sequences is assembled in the ring (byte stores
//           never leave the SM), matches whose source is younger than ~3 KiB read the ring, older ones
//           read global memory (flushed long ago: never on the dependency path), and the ring is flushed
//           in whole 512-byte chunks with aligned 16-byte stores (full sectors, no read-modify-write).
//
// Block checksums: G <= 8 blocks per warp, hashed concurrently by the eight quads before the decode
// (Check_Checksum, lib/lz4ada.adb:698-707), then decoded one after the other.
// Anything unusual (all error conditions, matches reaching before the block) goes to the exact
// routine process_block, which owns the reference's error semantics (lib/lz4ada.adb:716-904).
#pragma once

#include "kernels_v2.cuh"

namespace lz4b200 {
namespace v4 {

constexpr int WARPS = 4;                     // per CTA
constexpr uint32_t SEG = 132;                // 33 words: lane j's segment starts in bank j
constexpr uint32_t WIN = 32 * SEG;           // compressed window per warp
constexpr uint32_t CW_BYTES = WIN + 48;
constexpr uint32_t NTOK = 1024;              // token table per window (a window has <= 1408 tokens; text ~900)
constexpr uint32_t RING = 4096;
constexpr uint32_t BATCH_MAX = 1024;         // output bytes of a batch that goes through the ring
constexpr uint32_t CHUNK = 512;              // flush granule
constexpr uint32_t LIT_LANE = 16;            // literal runs up to this length: lane per sequence

enum : uint32_t { W_OK = 0, W_CUT = 1, W_BAD = 2 };

struct __align__(16) WarpMem {
	uint8_t ring[RING];
	uint8_t cw[CW_BYTES];
	uint16_t tokpos[NTOK];
};
static_assert(sizeof(WarpMem) % 16 == 0, "");

struct Walk { uint32_t x, c, o, st; };

// Token chain from window position p until it leaves [.., seg_hi), by every enabled lane at once
// (Decompress_Sequence, lib/lz4ada.adb:737-777; lengths: Process_Variable_Length :724-735).
// The loop is warp-uniform; a lane that is done idles through the remaining iterations.
// A sequence that needs bytes beyond the window stops the walk in front of its token: W_CUT (the
// next window starts there), or W_BAD when the window is the block's last (truncated block).
template <bool EMIT>
__device__ __forceinline__ Walk walk_segment(const uint8_t *cw, uint32_t p, uint32_t seg_hi, uint32_t wlen, bool last,
					     bool enable, uint32_t idx, uint16_t *tokpos)
{
	Walk r;
	r.c = 0; r.o = 0; r.st = W_OK;
	bool active = enable && p < seg_hi;
	while (__any_sync(FULL_MASK, active)) {
		if (active) {
			const uint32_t tk = cw[p];
			uint32_t lit = tk >> 4, ml = tk & 15;
			uint32_t nx = p + 3 + lit;
			bool okay = true;
			if (lit == 15 || ml == 15 || nx > wlen) {
				// out of line: extensions, the final sequence, the window edge
				uint32_t q = p + 1;
				bool cut = false;
				if (lit == 15) {
					uint32_t b;
					do {
						if (q >= wlen) { cut = true; break; }
						b = cw[q++];
						lit += b;
					} while (b == 255);
				}
				const uint32_t e = q + lit;
				nx = e;
				if (!cut) {
					if (e > wlen) {
						cut = true;
					} else if (last && e == wlen) {
						if (ml) { r.st = W_BAD; okay = false; }   // :752-764
						ml = 0;
					} else if (e + 2 > wlen) {
						cut = true;
					} else {
						nx = e + 2;
						if (ml == 15) {
							uint32_t b;
							do {
								if (nx >= wlen) { cut = true; break; }
								b = cw[nx++];
								ml += b;
							} while (b == 255);
						}
						ml += 4;
					}
				}
				if (cut) { r.st = last ? W_BAD : W_CUT; okay = false; }
			} else {
				ml += 4;
			}
			if (okay) {
				if (EMIT) {
					tokpos[idx] = static_cast<uint16_t>(p);
					idx++;
				}
				r.c++;
				r.o += lit + ml;
				p = nx;
			}
			active = okay && p < seg_hi;
		}
	}
	r.x = p;
	return r;
}

__device__ __forceinline__ uint32_t ridx(uint32_t x, uint32_t phase) { return (x + phase) & (RING - 1); }

// Ring -> global: output positions [from, to), 16-byte aligned on both sides wherever possible
// (ring index and global address share the same 16-byte phase).
__device__ __forceinline__ void flush_ring(const uint8_t *ring, uint32_t phase, uint8_t *og, uint32_t from, uint32_t to, int lane)
{
	if (from >= to) return;
	uint32_t a = from;
	const uint32_t head = (16u - ((from + phase) & 15u)) & 15u;
	const uint32_t h = head < to - from ? head : to - from;
	if (static_cast<uint32_t>(lane) < h) og[a + lane] = ring[ridx(a + lane, phase)];
	a += h;
	const uint32_t nvec = (to - a) >> 4;
	for (uint32_t v = lane; v < nvec; v += 32)
		*reinterpret_cast<uint4 *>(og + a + v * 16) = *reinterpret_cast<const uint4 *>(ring + ridx(a + v * 16, phase));
	a += nvec << 4;
	if (static_cast<uint32_t>(lane) < to - a) og[a + lane] = ring[ridx(a + lane, phase)];
}

// Byte i of output position x: wherever it lives (ring if young enough, else global).
__device__ __forceinline__ uint8_t out_byte(const uint8_t *ring, uint32_t phase, const uint8_t *og, uint32_t x, uint32_t near_lo)
{
	return x >= near_lo ? ring[ridx(x, phase)] : og[x];
}

// One match by the whole warp, destination in the ring, any length up to BATCH_MAX, any overlap
// (Output_With_History phases I and R, lib/lz4ada.adb:876-903).  Sources below near_lo are flushed.
__device__ __forceinline__ void coop_match(uint8_t *ring, uint32_t phase, const uint8_t *og, uint32_t mo, uint32_t off,
					   uint32_t ml, uint32_t near_lo, int lane)
{
	if (off >= ml) {
		for (uint32_t i = lane; i < ml; i += 32) ring[ridx(mo + i, phase)] = out_byte(ring, phase, og, mo - off + i, near_lo);
		return;
	}
	uint32_t done = 0;
	if (off < 32) {
		const uint32_t m0 = ml < 32 ? ml : 32;
		if (static_cast<uint32_t>(lane) < m0)
			ring[ridx(mo + lane, phase)] = out_byte(ring, phase, og, mo - off + (static_cast<uint32_t>(lane) % off), near_lo);
		done = m0;
		__syncwarp();
	}
	while (done < ml) {
		const uint32_t avail = off + done;
		const uint32_t L = avail - (avail % off);   // whole periods back: that source is valid
		const uint32_t chunk = (ml - done) < L ? (ml - done) : L;
		for (uint32_t i = lane; i < chunk; i += 32)
			ring[ridx(mo + done + i, phase)] = out_byte(ring, phase, og, mo + done - L + i, near_lo);
		done += chunk;
		__syncwarp();
	}
}

// Short non-overlapping match by one lane (ml <= 32, off >= ml): three aligned 16-byte loads from
// the ring (young source) or from global memory (old source), a byte shift that brings the match
// data to byte 0, a second one to the destination's word phase, then at most 3 head bytes + 8 aligned
// words + 3 tail bytes of stores into the ring.  maxml = warp-uniform bound on ml among the callers.
__device__ __forceinline__ void copy_simple(uint8_t *ring, uint32_t phase, const uint8_t *og, uint32_t src_s, uint32_t mo,
					    uint32_t ml, uint32_t maxml, bool is_far, bool active)
{
	if (!active) return;
	uint4 A, B = make_uint4(0, 0, 0, 0), C = make_uint4(0, 0, 0, 0);
	uint32_t m;
	if (is_far) {
		const uint8_t *sp = og + src_s;
		m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(sp) & 15u);
		const uint4 *base = reinterpret_cast<const uint4 *>(sp - m);
		A = base[0];
		if (m + ml > 16) B = base[1];
		if (m + ml > 32) C = base[2];
	} else {
		const uint32_t r0 = ridx(src_s, phase);
		m = r0 & 15u;
		const uint32_t b0 = r0 - m;
		A = *reinterpret_cast<const uint4 *>(ring + b0);
		if (m + ml > 16) B = *reinterpret_cast<const uint4 *>(ring + ((b0 + 16) & (RING - 1)));
		if (m + ml > 32) C = *reinterpret_cast<const uint4 *>(ring + ((b0 + 32) & (RING - 1)));
	}
	unsigned long long d0 = A.x | (static_cast<unsigned long long>(A.y) << 32);
	unsigned long long d1 = A.z | (static_cast<unsigned long long>(A.w) << 32);
	unsigned long long d2 = B.x | (static_cast<unsigned long long>(B.y) << 32);
	unsigned long long d3 = B.z | (static_cast<unsigned long long>(B.w) << 32);
	unsigned long long d4 = C.x | (static_cast<unsigned long long>(C.y) << 32);
	const unsigned long long d5 = C.z | (static_cast<unsigned long long>(C.w) << 32);
	if (m & 8) { d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = d5; }
	const uint32_t sh = (m & 7) * 8;
	if (sh) {
		d0 = (d0 >> sh) | (d1 << (64 - sh));
		d1 = (d1 >> sh) | (d2 << (64 - sh));
		d2 = (d2 >> sh) | (d3 << (64 - sh));
		d3 = (d3 >> sh) | (d4 << (64 - sh));
	}
	// match byte k is now byte k of the stream w0..w7
	const uint32_t w0 = static_cast<uint32_t>(d0), w1 = static_cast<uint32_t>(d0 >> 32), w2 = static_cast<uint32_t>(d1),
		       w3 = static_cast<uint32_t>(d1 >> 32), w4 = static_cast<uint32_t>(d2), w5 = static_cast<uint32_t>(d2 >> 32),
		       w6 = static_cast<uint32_t>(d3), w7 = static_cast<uint32_t>(d3 >> 32);
	const uint32_t a = ridx(mo, phase);
	if (a + ml > RING) {
		// the destination wraps around the end of the ring (about one lane in a hundred): masked byte stores
		const uint32_t ws[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
#pragma unroll
		for (int k = 0; k < 32; k++)
			if (static_cast<uint32_t>(k) < ml) ring[(a + k) & (RING - 1)] = static_cast<uint8_t>(ws[k >> 2] >> (8 * (k & 3)));
		return;
	}
	uint8_t *dp = ring + a;
	const uint32_t hb0 = (4u - (a & 3u)) & 3u;
	const uint32_t hb = hb0 < ml ? hb0 : ml;   // head bytes up to the next word boundary
	if (hb > 0) dp[0] = static_cast<uint8_t>(w0);
	if (hb > 1) dp[1] = static_cast<uint8_t>(w0 >> 8);
	if (hb > 2) dp[2] = static_cast<uint8_t>(w0 >> 16);
	const uint32_t hs = hb * 8;
	const uint32_t nwords = (ml - hb) >> 2;
	uint32_t *dw = reinterpret_cast<uint32_t *>(dp + hb);
	// v[j] = match bytes [hb + 4j, hb + 4j + 4)
	const uint32_t v0 = __funnelshift_r(w0, w1, hs), v1 = __funnelshift_r(w1, w2, hs);
	uint32_t v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0, v7 = 0;
	if (0 < nwords) dw[0] = v0;
	if (1 < nwords) dw[1] = v1;
	if (maxml > 8) {
		v2 = __funnelshift_r(w2, w3, hs);
		v3 = __funnelshift_r(w3, w4, hs);
		if (2 < nwords) dw[2] = v2;
		if (3 < nwords) dw[3] = v3;
	}
	if (maxml > 16) {
		v4 = __funnelshift_r(w4, w5, hs);
		v5 = __funnelshift_r(w5, w6, hs);
		if (4 < nwords) dw[4] = v4;
		if (5 < nwords) dw[5] = v5;
	}
	if (maxml > 24) {
		v6 = __funnelshift_r(w6, w7, hs);
		v7 = __funnelshift_r(w7, 0u, hs);
		if (6 < nwords) dw[6] = v6;
		if (7 < nwords) dw[7] = v7;
	}
	// tail: bytes hb + 4 * nwords .. ml - 1 (at most 3) are the low bytes of v[nwords]
	const uint32_t tb = hb + 4u * nwords;
	if (tb < ml) {
		const uint32_t t = (nwords & 4) ? ((nwords & 2) ? ((nwords & 1) ? v7 : v6) : ((nwords & 1) ? v5 : v4))
						: ((nwords & 2) ? ((nwords & 1) ? v3 : v2) : ((nwords & 1) ? v1 : v0));
		dp[tb] = static_cast<uint8_t>(t);
		if (tb + 1 < ml) dp[tb + 1] = static_cast<uint8_t>(t >> 8);
		if (tb + 2 < ml) dp[tb + 2] = static_cast<uint8_t>(t >> 16);
	}
}

struct BlockState {
	uint32_t pos;        // output bytes produced (block-relative)
	uint32_t flushed;    // output bytes already in global memory
	uint32_t ring_lo;    // lowest output position the ring is known to hold
};

// A batch whose output does not fit the ring protocol (a long match or literal run): bring global
// memory up to date, do the batch sequence by sequence straight in global memory with the v1 warp
// copies, then re-seed the ring with the last KiB so that the following batches find their near history.
__device__ __noinline__ void slow_batch(uint8_t *ring, uint32_t phase, uint8_t *og, const uint8_t *cwg, BlockState &st,
					uint32_t cnt, uint32_t lit, uint32_t ml, uint32_t off, uint32_t q, uint32_t total, int lane)
{
	flush_ring(ring, phase, og, st.flushed, st.pos, lane);
	__syncwarp();
	uint32_t p = st.pos;
	for (uint32_t j = 0; j < cnt; j++) {
		const uint32_t litj = __shfl_sync(FULL_MASK, lit, j), mlj = __shfl_sync(FULL_MASK, ml, j);
		const uint32_t offj = __shfl_sync(FULL_MASK, off, j), qj = __shfl_sync(FULL_MASK, q, j);
		if (litj) warp_copy<true>(og + p, cwg + qj, litj, lane);   // literals straight from the compressed stream
		p += litj;
		if (mlj) match_copy(og + p, offj, mlj, lane);
		p += mlj;
		__syncwarp();
	}
	st.pos += total;
	st.flushed = st.pos;
	const uint32_t keep = st.pos < BATCH_MAX ? st.pos : BATCH_MAX;
	st.ring_lo = st.pos - keep;
	for (uint32_t i = lane; i < keep; i += 32) ring[ridx(st.ring_lo + i, phase)] = og[st.ring_lo + i];
	__syncwarp();
}

// One batch: sequences 32k .. 32k+31 of the window's token table, lane per sequence.
// Returns false when the block needs the exact routine (offset 0, match reaching before the block).
__device__ __forceinline__ bool batch(WarpMem &wm, uint32_t phase, uint8_t *og, const uint8_t *cw, const uint8_t *cwg,
				      uint32_t k, uint32_t T, uint32_t wlen, bool last, BlockState &st, int lane)
{
	uint8_t *ring = wm.ring;
	const uint32_t idx = k * 32 + lane;
	const bool act = idx < T;
	const uint32_t cnt = T - k * 32 < 32 ? T - k * 32 : 32;
	uint32_t lit = 0, ml = 0, off = 0, q = 0;
	if (act) {
		const uint32_t t = wm.tokpos[idx];
		const uint32_t tk = cw[t];
		lit = tk >> 4;
		q = t + 1;
		if (lit == 15) {
			uint32_t b;
			do {
				b = cw[q++];
				lit += b;
			} while (b == 255);
		}
		const uint32_t e = q + lit;
		if (!(last && e == wlen)) {
			off = cw[e] | (static_cast<uint32_t>(cw[e + 1]) << 8);
			ml = tk & 15;
			if (ml == 15) {
				uint32_t nx = e + 2, b;
				do {
					b = cw[nx++];
					ml += b;
				} while (b == 255);
			}
			ml += 4;
		}
	}
	__syncwarp();
	const uint32_t len = lit + ml;
	uint32_t incl = len;
#pragma unroll
	for (int s = 1; s < 32; s <<= 1) {
		const uint32_t v = __shfl_up_sync(FULL_MASK, incl, s);
		if (lane >= s) incl += v;
	}
	const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
	const uint32_t B0 = st.pos;
	const uint32_t out_pos = B0 + incl - len;
	const uint32_t mo = out_pos + lit;
	if (__any_sync(FULL_MASK, ml && (off == 0 || off > mo))) return false;
	if (total > BATCH_MAX) {
		slow_batch(ring, phase, og, cwg, st, cnt, lit, ml, off, q, total, lane);
		return true;
	}
	// what the ring holds while this batch is being written
	const uint32_t lo_wr = B0 + total > RING ? B0 + total - RING : 0u;
	const uint32_t near_lo = st.ring_lo > lo_wr ? st.ring_lo : lo_wr;

	// ---- literals: lane per sequence up to LIT_LANE bytes, longer runs by the whole warp ----
	const uint32_t maxlit = __reduce_max_sync(FULL_MASK, lit);
	if (maxlit) {
		const uint32_t lim = maxlit < LIT_LANE ? maxlit : LIT_LANE;
		const uint32_t d = out_pos + phase;
		for (uint32_t i = 0; i < lim; i++)
			if (i < lit) ring[(d + i) & (RING - 1)] = cw[q + i];
		if (maxlit > LIT_LANE) {
			uint32_t big = __ballot_sync(FULL_MASK, lit > LIT_LANE);
			while (big) {
				const int j = __ffs(big) - 1;
				big &= big - 1;
				const uint32_t dj = __shfl_sync(FULL_MASK, out_pos, j) + phase, qj = __shfl_sync(FULL_MASK, q, j);
				const uint32_t lj = __shfl_sync(FULL_MASK, lit, j);
				for (uint32_t i = LIT_LANE + lane; i < lj; i += 32) ring[(dj + i) & (RING - 1)] = cw[qj + i];
			}
		}
	}

	// ---- matches ----
	const uint32_t src_s = mo - off;
	const uint32_t src_e = src_s + (ml < off ? ml : off);   // self-overlap: the source ends where the match starts
	const bool simple = ml <= 32 && off >= ml;
	const bool is_far = src_s < near_lo;                     // then the whole source is in global memory
	bool done = (ml == 0);
	// dep = earlier sequences of this batch whose output overlaps my source
	uint32_t dep = 0;
	if (__any_sync(FULL_MASK, !done && src_e > B0)) {
		uint32_t lo = 0, hi = 0;
		const uint32_t out_end = out_pos + len;
#pragma unroll
		for (int step = 16; step >= 1; step >>= 1) {
			const uint32_t e = __shfl_sync(FULL_MASK, out_end, (lo + step - 1) & 31);
			const uint32_t b = __shfl_sync(FULL_MASK, out_pos, (hi + step - 1) & 31);
			if (lo + step <= 32 && e <= src_s) lo += step;
			if (hi + step <= 32 && b < src_e) hi += step;
		}
		const uint32_t below_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
		const uint32_t below_lo = lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u);
		dep = below_hi & ~below_lo & ((1u << lane) - 1u);
		if (done || src_e <= B0) dep = 0;
	}
	__syncwarp();
	// round A: every match that waits for nothing inside the batch -- short ones lane per sequence, all at once
	{
		const bool ready = !done && dep == 0;
		const bool rs = ready && simple;
		if (__any_sync(FULL_MASK, rs)) {
			const uint32_t maxml = __reduce_max_sync(FULL_MASK, rs ? ml : 0u);
			copy_simple(ring, phase, og, src_s, mo, ml, maxml, is_far, rs);
		}
		done = done || rs;
	}
	// the rest strictly in stream order, one match at a time by the whole warp: matches that wait for
	// output of this batch (source in the ring, a few cycles away), long and self-overlapping ones.
	// One packed word per lane serves the common case with a single shuffle.
	uint32_t rest = __ballot_sync(FULL_MASK, !done);
	if (rest) {
		const bool quick = !done && simple && !is_far;
		const uint32_t packed = quick ? (0x80000000u | ((mo - B0) << 16) | ((ml - 1) << 11) | (off & 0x7ffu)) : 0u;
		const bool quick2 = quick && off < 0x800u;
		const uint32_t pk_mine = quick2 ? packed : 0u;
		while (rest) {
			const int j = __ffs(rest) - 1;
			rest &= rest - 1;
			const uint32_t pk = __shfl_sync(FULL_MASK, pk_mine, j);
			__syncwarp();
			if (pk & 0x80000000u) {
				const uint32_t moj = B0 + ((pk >> 16) & 0x7fffu), mlj = ((pk >> 11) & 31u) + 1u, offj = pk & 0x7ffu;
				if (static_cast<uint32_t>(lane) < mlj) ring[ridx(moj + lane, phase)] = ring[ridx(moj - offj + lane, phase)];
				continue;
			}
			coop_match(ring, phase, og, __shfl_sync(FULL_MASK, mo, j), __shfl_sync(FULL_MASK, off, j),
				   __shfl_sync(FULL_MASK, ml, j), near_lo, lane);
		}
		__syncwarp();
	}
	st.pos = B0 + total;
	// ---- flush whole chunks ----
	const uint32_t target = ((st.pos + phase) & ~(CHUNK - 1));
	if (target > st.flushed + phase) {
		flush_ring(ring, phase, og, st.flushed, target - phase, lane);
		st.flushed = target - phase;
		__syncwarp();
	}
	return true;
}

// One compressed block, start to finish, by one warp.  false = give it to the exact routine.
__device__ __forceinline__ bool decode_block(const uint8_t *__restrict__ s, uint32_t n, uint8_t *og, uint32_t cap,
					     WarpMem &wm, int lane, uint32_t &out_len)
{
	const uint32_t phase = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(og) & 15u);
	BlockState st = {0, 0, 0};
	uint32_t ip = 0;
	while (ip < n) {
		// ---------------- load the window ----------------
		const uint8_t *g0 = s + ip;
		const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g0) & 15u);
		const uint32_t wlen = n - ip < WIN ? n - ip : WIN;
		const bool last = ip + wlen == n;
		const uint32_t nvec = (mis + wlen + 15u) >> 4;
		__syncwarp();
		for (uint32_t v = lane; v < nvec; v += 32) cp_async16(wm.cw + v * 16, g0 - mis + v * 16);
		cp_async_commit();
		cp_async_wait<0>();
		__syncwarp();
		const uint8_t *cw = wm.cw + mis;

		// ---------------- parse: fixed point of "my entry = my left neighbour's exit" ----------------
		const uint32_t seg_lo = static_cast<uint32_t>(lane) * SEG;
		const uint32_t seg_hi = seg_lo + SEG < wlen ? seg_lo + SEG : wlen;
		uint32_t g = seg_lo < wlen ? seg_lo : wlen;
		Walk r = walk_segment<false>(cw, g, seg_hi, wlen, last, true, 0, nullptr);
		for (;;) {
			uint32_t ng = __shfl_up_sync(FULL_MASK, r.x, 1);
			if (lane == 0) ng = 0;
			const bool changed = ng != g;
			if (!__any_sync(FULL_MASK, changed)) break;
			if (changed) g = ng;
			const Walk r2 = walk_segment<false>(cw, g, seg_hi, wlen, last, changed, 0, nullptr);
			if (changed) r = r2;
		}
		if (__any_sync(FULL_MASK, r.st == W_BAD)) return false;
		// ---------------- scan: sequences and output bytes in front of every segment ----------------
		uint32_t ic = r.c, io = r.o;
#pragma unroll
		for (int sft = 1; sft < 32; sft <<= 1) {
			const uint32_t a = __shfl_up_sync(FULL_MASK, ic, sft), c2 = __shfl_up_sync(FULL_MASK, io, sft);
			if (lane >= sft) { ic += a; io += c2; }
		}
		const bool use = ic <= NTOK;   // leading segments that fit the token table
		const uint32_t nuse = __popc(__ballot_sync(FULL_MASK, use));
		if (nuse == 0) return false;
		const uint32_t T = __shfl_sync(FULL_MASK, ic, nuse - 1), O = __shfl_sync(FULL_MASK, io, nuse - 1);
		const uint32_t wend = __shfl_sync(FULL_MASK, r.x, nuse - 1);
		if (wend == 0 || T == 0 || O > cap - st.pos) return false;
		// ---------------- emit the token table ----------------
		walk_segment<true>(cw, g, seg_hi, wlen, last, use && r.c != 0, ic - r.c, wm.tokpos);
		__syncwarp();
		// ---------------- batches ----------------
		const uint32_t nb = (T + 31) >> 5;
		for (uint32_t k = 0; k < nb; k++)
			if (!batch(wm, phase, og, cw, g0, k, T, wlen, last, st, lane)) return false;
		ip += wend;
	}
	flush_ring(wm.ring, phase, og, st.flushed, st.pos, lane);
	out_len = st.pos;
	return true;
}

// One warp, G <= 8 blocks (first_block .. first_block + G - 1): hash together, decode in turn.
__device__ __forceinline__ void decode_group(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks, uint32_t first_block,
					     uint32_t G, const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status,
					     WarpMem &wm, int lane)
{
	uint32_t state = LS_IDLE;
	const uint8_t *s = src;
	uint8_t *o = dst;
	uint32_t n = 0, cap = 0, flags = 0, op = 0;
	uint32_t computed = 0, declared = 0, code = LZ4B200_ST_OK;
	if (static_cast<uint32_t>(lane) < G && first_block + lane < n_blocks) {
		const lz4b200_blk_desc d = desc[first_block + lane];
		flags = d.flags;
		if (!(flags & LZ4B200_BLK_CHAINED)) {
			state = LS_RUN;
			s = src + d.src_off;
			o = dst + d.dst_off;
			n = d.src_len;
			cap = d.dst_cap;
		}
	}
	// ---- fused block checksum: quad q hashes block q (lib/lz4ada.adb:698-707) ----
	{
		const int qd = lane >> 2;
		const uint8_t *sq = shfl_cptr(s, qd);
		const uint32_t nq = __shfl_sync(FULL_MASK, n, qd);
		const uint32_t fq = __shfl_sync(FULL_MASK, flags, qd);
		const uint32_t stq = __shfl_sync(FULL_MASK, state, qd);
		const bool want = stq == LS_RUN && (fq & LZ4B200_BLK_HAS_CHECKSUM);
		if (__any_sync(FULL_MASK, want)) {
			const uint32_t h = quad_xxh32_prologue(sq, want ? nq : 0, lane);
			const uint32_t hq = __shfl_sync(FULL_MASK, h, (lane * 4) & 31);
			if (lane < 8 && state == LS_RUN && (flags & LZ4B200_BLK_HAS_CHECKSUM)) {
				const uint8_t *t = s + n;
				declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
				computed = hq;
				if (computed != declared) {
					code = LZ4B200_ST_BLOCK_CHECKSUM;
					state = LS_DONE;
				}
			}
		}
	}
	if (state == LS_RUN && (flags & LZ4B200_BLK_HASH_ONLY)) state = LS_DONE;
#pragma unroll 1
	for (uint32_t g = 0; g < G; g++) {
		const uint32_t stg = __shfl_sync(FULL_MASK, state, g);
		if (stg != LS_RUN) continue;
		const uint32_t fg = __shfl_sync(FULL_MASK, flags, g);
		const uint32_t ng = __shfl_sync(FULL_MASK, n, g);
		const uint32_t capg = __shfl_sync(FULL_MASK, cap, g);
		const uint8_t *sg = shfl_cptr(s, g);
		uint8_t *og = shfl_ptr(o, g);
		if (fg & LZ4B200_BLK_STORED) {   // lib/lz4ada.adb:685-695
			if (ng > capg) {
				if (static_cast<uint32_t>(lane) == g) state = LS_FALLBACK;
				continue;
			}
			warp_copy<true>(og, sg, ng, lane);
			if (static_cast<uint32_t>(lane) == g) {
				op = ng;
				state = LS_DONE;
			}
			continue;
		}
		uint32_t produced = 0;
		const bool okay = decode_block(sg, ng, og, capg, wm, lane, produced);
		__syncwarp();
		if (static_cast<uint32_t>(lane) == g) {
			if (okay) {
				op = produced;
				state = LS_DONE;
			} else {
				state = LS_FALLBACK;
			}
		}
	}
	if (state == LS_DONE) {
		lz4b200_blk_status *st = status + first_block + lane;
		st->code = code;
		st->out_len = code == LZ4B200_ST_OK ? op : 0;
		st->err_pos = 0;
		st->aux = 0;
		st->xxh32_computed = computed;
		st->xxh32_declared = declared;
	}
	// ---- exact path for whatever the fast path gave up on ----
#pragma unroll 1
	for (uint32_t g = 0; g < G; g++) {
		if (__shfl_sync(FULL_MASK, state, g) != LS_FALLBACK) continue;
		const uint32_t b = first_block + g;
		const lz4b200_blk_desc d = desc[b];
		process_block<false>(src, dst + d.dst_off, d, d.dst_cap, d.hist_avail, status + b, lane);
	}
}

}  // namespace v4
}  // namespace lz4b200
