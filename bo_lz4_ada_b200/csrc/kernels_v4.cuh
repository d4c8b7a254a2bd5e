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

constexpr int WARPS = 4;                     // per CTA at most (the launch uses one warp per CTA for batches of few, long-running blocks)
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

struct Walk { uint32_t x, c, st; };

constexpr uint32_t STOP = 0xffffffffu;

// One token that is not the plain case (a 15 nibble, the final sequence, the window edge), out of line.
// Returns the position of the next token, or STOP with st = W_CUT / W_BAD when the walk ends in front
// of this token (Decompress_Sequence, lib/lz4ada.adb:737-777; lengths: Process_Variable_Length :724-735).
__device__ __noinline__ uint32_t token_slow(const uint8_t *cw, uint32_t p, uint32_t wlen, uint32_t last, uint32_t &st)
{
	// A length with more than GIANT_EXT extension bytes (a run of 256 KiB and more: zero pages, RLE) ends the window
	// in front of its token (W_CUT): decode_block takes such a sequence on its own -- the whole warp scans the length
	// 32 bytes at a time and copies in global memory -- while the speculative chains here would crawl through the
	// 0xff bytes one by one.
	constexpr uint32_t GIANT_EXT = 1024;
	const uint32_t tk = cw[p];
	uint32_t lit = tk >> 4, q = p + 1;
	const uint32_t stop_st = last ? W_BAD : W_CUT;
	if (lit == 15) {
		uint32_t b;
		const uint32_t q0 = q;
		do {
			if (q >= wlen) { st = stop_st; return STOP; }
			if (q - q0 > GIANT_EXT) { st = W_CUT; return STOP; }
			b = cw[q++];
			lit += b;
		} while (b == 255);
	}
	const uint32_t e = q + lit;   // end of the literals
	if (e > wlen) { st = stop_st; return STOP; }
	if (last && e == wlen) {
		// final literal-only sequence (:752-764); a match nibble here is an error
		if (tk & 15) { st = W_BAD; return STOP; }
		return e;
	}
	if (e + 2 > wlen) { st = stop_st; return STOP; }
	uint32_t nx = e + 2;
	if ((tk & 15) == 15) {
		uint32_t b;
		const uint32_t n0 = nx;
		do {
			if (nx >= wlen) { st = stop_st; return STOP; }
			if (nx - n0 > GIANT_EXT) { st = W_CUT; return STOP; }
			b = cw[nx++];
		} while (b == 255);
	}
	return nx;
}

// Position of the token after the one at p (STOP + st as token_slow).  In line: no extension bytes or a
// single one per length (literal runs < 270, matches < 274); the rest goes out of line.  Garbage chains
// of the speculative parse read "tokens" like 'o' = 0x6f all the time, so the 15 nibbles must be cheap.
__device__ __forceinline__ uint32_t token_next(const uint8_t *cw, uint32_t p, uint32_t wlen, uint32_t last, uint32_t &st)
{
	const uint32_t tk = cw[p];
	uint32_t lit = tk >> 4, q = p + 1;
	bool slow = false;
	if (lit == 15) {
		// q < wlen + 1 <= CW_BYTES: the read is inside the buffer even at the window edge
		const uint32_t b = cw[q];
		q++;
		lit += b;
		slow = b == 255;
	}
	uint32_t nx = q + lit + 2;
	if ((tk & 15) == 15 && nx < wlen) {
		const uint32_t b = cw[nx];
		nx++;
		slow = slow || b == 255;
	}
	// nx >= wlen: the window edge, the final sequence, or a 15 nibble whose extension byte was not read
	if (slow || nx >= wlen) return token_slow(cw, p, wlen, last, st);
	return nx;
}

// Token chain from window position p until it leaves [.., seg_hi), by every enabled lane at once.
// The loop is warp-uniform; a lane that is done idles through the remaining iterations.
// A sequence that needs bytes beyond the window stops the walk in front of its token: W_CUT (the
// next window starts there), or W_BAD when the window is the block's last (truncated block).
// EMIT: record the token positions from index idx on.
template <bool EMIT>
__device__ __forceinline__ Walk walk_segment(const uint8_t *cw, uint32_t p, uint32_t seg_hi, uint32_t wlen, uint32_t last,
					     bool enable, uint32_t idx, uint16_t *tokpos)
{
	Walk r;
	r.c = 0; r.st = W_OK;
	uint32_t end = enable ? seg_hi : 0u;   // the lane is active while p < end
	while (__any_sync(FULL_MASK, p < end)) {
		if (p < end) {
			uint32_t nx = token_next(cw, p, wlen, last, r.st);
			if (nx != STOP) {
				if (EMIT) tokpos[idx++] = static_cast<uint16_t>(p);
				r.c++;
				p = nx;
			} else {
				end = 0;
			}
		}
	}
	r.x = p;
	return r;
}

// Re-walk after the entry of a segment moved from g_old to g_new, for the lanes with `changed`: the
// new chain normally joins the old one after a few tokens, and from there on exit, status and count
// are the old ones.  Both chains advance in lock-step (always the one that is behind) until they
// stand on the same position or have both left the segment.
__device__ __forceinline__ void rewalk_merge(const uint8_t *cw, uint32_t g_new, uint32_t g_old, uint32_t seg_hi, uint32_t wlen,
					     uint32_t last, bool changed, Walk &r)
{
	uint32_t a = g_new, b = g_old, ca = 0, cb = 0, st_a = W_OK;
	// the old chain's tokens all lie below b_end; it ends standing on r.x
	const uint32_t b_end = r.x < seg_hi ? r.x : seg_hi;
	bool run = changed;
	while (__any_sync(FULL_MASK, run)) {
		if (run) {
			const bool fin_a = a >= seg_hi || st_a != W_OK, fin_b = b >= b_end;
			if (a == b) {
				r.c = r.c - cb + ca;   // merged: exit and status stay
				run = false;
			} else if (fin_a && (fin_b || b > a)) {
				// no common token: the new chain stands on its own
				r.x = a; r.c = ca; r.st = st_a;
				run = false;
			} else {
				const bool step_a = !fin_a && (a < b || fin_b);
				const uint32_t t = step_a ? a : b;
				uint32_t st_t = W_OK;
				const uint32_t nx = token_next(cw, t, wlen, last, st_t);
				if (step_a) {
					if (nx == STOP) st_a = st_t;
					else { a = nx; ca++; }
				} else {
					// the old chain parsed this token before: it cannot stop here
					b = nx; cb++;
				}
			}
		}
	}
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
// the ring (young source) or from global memory (old source, issued early in the batch so that the L2
// round trip overlaps the literal copies), a byte shift that brings the match data to byte 0, a second
// one to the destination's word phase, then at most 3 head bytes + 8 aligned words + 3 tail bytes of
// stores into the ring.
struct Src48 {
	uint4 A, B, C;
	uint32_t m;   // byte offset of the match data inside A
};

__device__ __forceinline__ void load_far(Src48 &S, const uint8_t *og, uint32_t src_s, uint32_t ml, bool active)
{
	S.A = S.B = S.C = make_uint4(0, 0, 0, 0);
	S.m = 0;
	if (!active) return;
	const uint8_t *sp = og + src_s;
	S.m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(sp) & 15u);
	const uint4 *base = reinterpret_cast<const uint4 *>(sp - S.m);
	S.A = __ldcg(base);
	if (S.m + ml > 16) S.B = __ldcg(base + 1);
	if (S.m + ml > 32) S.C = __ldcg(base + 2);
}

__device__ __forceinline__ void load_near(Src48 &S, const uint8_t *ring, uint32_t phase, uint32_t src_s, uint32_t ml, bool active)
{
	if (!active) return;
	const uint32_t r0 = ridx(src_s, phase);
	S.m = r0 & 15u;
	const uint32_t b0 = r0 - S.m;
	S.A = *reinterpret_cast<const uint4 *>(ring + b0);
	if (S.m + ml > 16) S.B = *reinterpret_cast<const uint4 *>(ring + ((b0 + 16) & (RING - 1)));
	if (S.m + ml > 32) S.C = *reinterpret_cast<const uint4 *>(ring + ((b0 + 32) & (RING - 1)));
}

// maxml = warp-uniform bound on ml among the calling lanes.
__device__ __forceinline__ void store_simple(uint8_t *ring, uint32_t phase, const Src48 &S, uint32_t mo, uint32_t ml,
					     uint32_t maxml, bool active)
{
	if (!active) return;
	const uint4 A = S.A, B = S.B, C = S.C;
	const uint32_t m = S.m;
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
				      uint32_t k, uint32_t T, uint32_t wlen, bool last, uint32_t cap, BlockState &st, int lane)
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
	if (__any_sync(FULL_MASK, ml && (off == 0 || off > mo)) || total > cap - B0) return false;
	if (total > BATCH_MAX) {
		slow_batch(ring, phase, og, cwg, st, cnt, lit, ml, off, q, total, lane);
		return true;
	}
	// what the ring holds while this batch is being written
	const uint32_t lo_wr = B0 + total > RING ? B0 + total - RING : 0u;
	const uint32_t near_lo = st.ring_lo > lo_wr ? st.ring_lo : lo_wr;

	// ---- matches, part 1: classify, and start the global loads of the old sources right away ----
	const uint32_t src_s = mo - off;
	const uint32_t src_e = src_s + (ml < off ? ml : off);   // self-overlap: the source ends where the match starts
	const bool is_far = ml != 0 && src_s < near_lo;          // then the whole source is in global memory ...
	// ... except right behind a block's history (decode_block with hist > 0), where a source may begin in the
	// history and end in bytes of this block that only the ring holds: the byte-wise cooperative copy sorts it out
	const bool straddle = is_far && src_e > st.flushed;
	const bool simple = ml <= 32 && off >= ml && !straddle;
	Src48 S;
	load_far(S, og, src_s, ml, is_far && simple);

	// ---- literals: lane per sequence up to LIT_LANE bytes, longer runs by the whole warp ----
	const uint32_t maxlit = __reduce_max_sync(FULL_MASK, lit);
	if (maxlit) {
		const uint32_t d = ridx(out_pos, phase);
		const uint8_t *sp = cw + q;
		const uint32_t short_lit = lit < LIT_LANE ? lit : LIT_LANE;
		if (!__any_sync(FULL_MASK, d + short_lit > RING)) {
			uint8_t *dp = ring + d;
#pragma unroll
			for (int i4 = 0; i4 < static_cast<int>(LIT_LANE); i4 += 4) {
				if (static_cast<uint32_t>(i4) < maxlit) {
#pragma unroll
					for (int i = i4; i < i4 + 4; i++)
						if (static_cast<uint32_t>(i) < short_lit) dp[i] = sp[i];
				}
			}
		} else {
			// some lane's literals cross the end of the ring: masked stores
			for (uint32_t i = 0; i < LIT_LANE && i < maxlit; i++)
				if (i < short_lit) ring[(d + i) & (RING - 1)] = sp[i];
		}
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

	// ---- matches, part 2 ----
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
	// parallel rounds, lane per sequence: first every short match that waits for nothing inside the batch,
	// then -- as long as enough of them are ready at once -- those whose sources have just been completed
	uint32_t undone = __ballot_sync(FULL_MASK, !done);
	for (int round = 0;; round++) {
		const bool rs = !done && simple && (dep & undone) == 0;
		const uint32_t nrs = __popc(__ballot_sync(FULL_MASK, rs));
		if (nrs == 0 || (round > 0 && nrs < 6)) break;
		const uint32_t maxml = __reduce_max_sync(FULL_MASK, rs ? ml : 0u);
		load_near(S, ring, phase, src_s, ml, rs && !is_far);
		store_simple(ring, phase, S, mo, ml, maxml, rs);
		done = done || rs;
		__syncwarp();
		undone = __ballot_sync(FULL_MASK, !done);
	}
	// the rest strictly in stream order, one match at a time by the whole warp: matches that wait for
	// output of this batch (source in the ring, a few cycles away), long and self-overlapping ones.
	// One packed word per lane serves the common case with a single shuffle; two matches per trip so
	// that the second shuffle is in flight while the first match is copied.
	uint32_t rest = undone;
	if (rest) {
		const bool quick = !done && simple && !is_far && off < 0x800u;
		const uint32_t pk_mine = quick ? (0x80000000u | ((mo - B0) << 16) | ((ml - 1) << 11) | off) : 0u;
		auto one = [&](uint32_t pk, int j) {
			__syncwarp();
			if (pk & 0x80000000u) {
				const uint32_t moj = B0 + ((pk >> 16) & 0x7fffu), mlj = ((pk >> 11) & 31u) + 1u, offj = pk & 0x7ffu;
				if (static_cast<uint32_t>(lane) < mlj) ring[ridx(moj + lane, phase)] = ring[ridx(moj - offj + lane, phase)];
			} else {
				coop_match(ring, phase, og, __shfl_sync(FULL_MASK, mo, j), __shfl_sync(FULL_MASK, off, j),
					   __shfl_sync(FULL_MASK, ml, j), near_lo, lane);
			}
		};
		while (rest) {
			const int j1 = __ffs(rest) - 1;
			rest &= rest - 1;
			const bool two = rest != 0;
			const int j2 = two ? __ffs(rest) - 1 : j1;
			rest &= rest - 1;
			const uint32_t pk1 = __shfl_sync(FULL_MASK, pk_mine, j1), pk2 = __shfl_sync(FULL_MASK, pk_mine, j2);
			one(pk1, j1);
			if (two) one(pk2, j2);
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
// The block's output starts at og + hist; the hist bytes in front of it are final history of the same frame
// that matches may reach into (0 for the independent blocks of K1; the stream window of the Update path).
// Positions inside are relative to og, so an old source is simply a global read below the ring.
__device__ __forceinline__ bool decode_block(const uint8_t *__restrict__ s, uint32_t n, uint8_t *og, uint32_t cap,
					     WarpMem &wm, int lane, uint32_t &out_len, uint32_t hist = 0)
{
	const uint32_t phase = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(og) & 15u);
	BlockState st = {hist, hist, hist};
	cap += hist;
	uint32_t ip = 0;
	while (ip < n) {
		// ---------------- load the window ----------------
		const uint8_t *g0 = s + ip;
		const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g0) & 15u);
		const uint32_t wlen = n - ip < WIN ? n - ip : WIN;
		const uint32_t last = ip + wlen == n ? 1u : 0u;
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
		for (uint32_t pass = 0;; pass++) {
			uint32_t ng = __shfl_up_sync(FULL_MASK, r.x, 1);
			if (lane == 0) ng = 0;
			const bool changed = ng != g;
			if (!__any_sync(FULL_MASK, changed)) break;
			if (pass == 0) {
				// nearly every lane moves from its guess to the true entry: plain walk
				const Walk r2 = walk_segment<false>(cw, ng, seg_hi, wlen, last, changed, 0, nullptr);
				if (changed) r = r2;
			} else {
				// a few lanes move again: follow the new chain only until it joins the old one
				rewalk_merge(cw, ng, g, seg_hi, wlen, last, changed, r);
			}
			if (changed) g = ng;
		}
		if (__any_sync(FULL_MASK, r.st == W_BAD)) return false;
		// ---------------- scan: sequences in front of every segment ----------------
		uint32_t ic = r.c;
#pragma unroll
		for (int sft = 1; sft < 32; sft <<= 1) {
			const uint32_t a = __shfl_up_sync(FULL_MASK, ic, sft);
			if (lane >= sft) ic += a;
		}
		const bool use = ic <= NTOK;   // leading segments that fit the token table
		const uint32_t nuse = __popc(__ballot_sync(FULL_MASK, use));
		if (nuse == 0) return false;
		const uint32_t T = __shfl_sync(FULL_MASK, ic, nuse - 1);
		const uint32_t wend = __shfl_sync(FULL_MASK, r.x, nuse - 1);
		if (wend == 0 || T == 0) {
			// the window's first token does not complete inside it (a literal run longer than the window, a length
			// field of hundreds of bytes): that one sequence straight in global memory, then on with the next window --
			// not the whole block to the exact routine, which walks a 450 KiB block for 18 ms
			uint32_t lp = 0, lit = 0, ml = 0, nxt = 0, off = 0;
			const bool fine = parse_token_wide(s, n, ip, lp, lit, ml, nxt, lane);
			if (fine && ml) off = ld_u8<true>(s + lp + lit) | (ld_u8<true>(s + lp + lit + 1) << 8);
			const uint32_t left = cap - st.pos;
			if (!fine || nxt <= ip || lit > left || ml > left - lit || (ml && (off == 0 || off > st.pos + lit))) return false;
			slow_batch(wm.ring, phase, og, s, st, 1, lit, ml, off, lp, lit + ml, lane);
			ip = nxt;
			continue;
		}
		// ---------------- emit the token table ----------------
		walk_segment<true>(cw, g, seg_hi, wlen, last, use && r.c != 0, ic - r.c, wm.tokpos);
		__syncwarp();
		// ---------------- batches ----------------
		const uint32_t nb = (T + 31) >> 5;
		for (uint32_t k = 0; k < nb; k++)
			if (!batch(wm, phase, og, cw, g0, k, T, wlen, last != 0, cap, st, lane)) return false;
		ip += wend;
	}
	flush_ring(wm.ring, phase, og, st.flushed, st.pos, lane);
	out_len = st.pos - hist;
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
		if (!(flags & LZ4B200_BLK_NOT_K1)) {
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
		// a block that opens with a match of 529 bytes or more (two 255 extension bytes: zero pages, RLE) is a
		// handful of giant sequences: the exact routine's warp-wide length scan and pattern replication are the
		// right tool, and the speculative parse would only chew on runs of 0xff
		bool giant_first = false;
		if (ng >= 8) {
			const uint32_t tk = ld_u8<true>(sg);
			uint32_t l = tk >> 4, q = 1;
			bool plain_len = true;
			if (l == 15) {
				const uint32_t e = ld_u8<true>(sg + 1);
				l += e;
				q = 2;
				plain_len = e != 255u;
			}
			const uint32_t x = q + l + 2;   // first extension byte of the match length
			if (plain_len && (tk & 15u) == 15u && x + 1 < ng) giant_first = ld_u8<true>(sg + x) == 255u && ld_u8<true>(sg + x + 1) == 255u;
		}
		uint32_t produced = 0;
		const bool okay = !giant_first && decode_block(sg, ng, og, capg, wm, lane, produced);
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
