// kernels_v2.cuh -- K1 fast path: G independent blocks per warp.
//
// v1 (kernels.cuh: one warp walks one block, every lane redundantly) measured ~90 instructions and
// ~1000 cycles of dependent latency per LZ4 sequence (profiles/r01_v1_*): token load -> literal
// load/store -> match load through L2 -> store, strictly one sequence at a time.  Text-like data has
// ~10 output bytes per sequence, so the path is bound by per-sequence latency and LSU issue, not HBM.
//
// v2 splits the work of a warp into two phases that each touch many sequences per instruction:
//
//   parse   lane g (< G) walks the token chain of its own block g and records up to 32 sequences
//           {literal position, literal length, match length} in shared memory.  The chain is
//           serial per block, so G chains advance per instruction instead of one.
//   copy    for each block in turn, lane s takes sequence s of the batch: all 32 offsets are loaded
//           by one instruction, output positions come from a warp scan, literals are copied lane
//           per sequence, matches whose source is already final are copied lane per sequence with
//           two aligned 16-byte loads + a byte shift (one wavefront per source line instead of one
//           per byte), and the few matches that depend on output of the same batch wait for the
//           "final frontier" to pass their source (rounds), or go through the warp-cooperative
//           copy when they are long or self-overlapping.
//
// Anything unusual -- any error condition, a match reaching before the block, lengths beyond the
// 16-bit descriptor fields -- abandons the fast path for that block and re-decodes it with the
// exact v1 routine (process_block), which owns the reference's error semantics.
#pragma once

#include "kernels.cuh"

namespace lz4b200 {

struct SeqDesc {          // 8 bytes, one per parsed sequence
	uint32_t lit_pos;     // block-relative position of the first literal byte
	uint16_t lit;         // literal length
	uint16_t ml;          // match length (0 = final literal-only sequence)
};

enum : uint32_t { LS_IDLE = 0, LS_RUN = 1, LS_DONE = 2, LS_FALLBACK = 3 };

__device__ __forceinline__ const uint8_t *shfl_cptr(const uint8_t *p, int srcLane)
{
	unsigned long long v = reinterpret_cast<unsigned long long>(p);
	v = __shfl_sync(FULL_MASK, v, srcLane);
	return reinterpret_cast<const uint8_t *>(v);
}
__device__ __forceinline__ uint8_t *shfl_ptr(uint8_t *p, int srcLane)
{
	unsigned long long v = reinterpret_cast<unsigned long long>(p);
	v = __shfl_sync(FULL_MASK, v, srcLane);
	return reinterpret_cast<uint8_t *>(v);
}

// Sequence with a 15 nibble: follow the extension bytes (Process_Variable_Length,
// lib/lz4ada.adb:724-735).  p = first byte after the token on entry, first literal byte on return.
// Returns false for anything the fast path does not take (truncation, fields beyond 16 bits).
__device__ __noinline__ bool parse_extended(const uint8_t *__restrict__ s, uint32_t n, uint32_t &p,
					    uint32_t &lit, uint32_t &ml, uint32_t &nxt)
{
	if (lit == 15) {
		uint32_t b;
		do {
			if (p >= n) return false;
			b = ld_u8<true>(s + p);
			p++;
			lit += b;
		} while (b == 255 && lit < 66000);
		if (lit > 65535) return false;
	}
	const uint32_t q = p + lit;
	if (q > n) return false;
	if (q == n) {
		if (ml) return false;
		nxt = n;
		return true;
	}
	if (q + 1 >= n) return false;
	nxt = q + 2;
	if (ml == 15) {
		uint32_t b;
		do {
			if (nxt >= n) return false;
			b = ld_u8<true>(s + nxt);
			nxt++;
			ml += b;
		} while (b == 255 && ml < 66000);
	}
	ml += 4;
	return ml <= 65535;
}

// The fused block checksum is a prologue; keep its register needs out of the decode loops.
__device__ __noinline__ uint32_t quad_xxh32_prologue(const uint8_t *p, uint32_t n, int lane)
{
	return quad_xxh32<true, false>(p, n, lane);
}

// Ordering gate for batches of one chain that are copied by different warps of a CTA (pipelined chain
// kernel): batch `my_batch` may read output before `out_start[d % slots]` where d = *done_upto (all
// batches < d are final); everything else of its past becomes final once *done_upto == my_batch.
struct BatchGate {
	volatile uint32_t *done_upto;
	const volatile uint32_t *out_start;   // ring: frame-relative output position where batch k starts
	const volatile uint32_t *base_lo, *base_hi;   // ring: chain-relative start of batch k's frame
	uint32_t my_base_lo, my_base_hi;
	volatile uint32_t *fail;
	uint32_t slots;
	uint32_t my_batch;
};

constexpr int SD_STRIDE = 33;           // descriptors per block row (32 used): odd stride, no bank conflicts
constexpr uint32_t TILE_BYTES = 1024;   // per-warp staging tile for one batch of output

// Copy one batch (c <= 32 sequences of one block).  Returns false when the batch needs the exact
// path (offset 0, match reaching before the block start, output capacity).  Warp-uniform result.
//
// The batch's output [opg, opg + total) is assembled in a shared-memory tile whenever it fits:
// byte-granular stores then never reach L2 (where partial-sector writes to evicted lines cost a
// DRAM read-modify-write each: 6 GB of DRAM reads per GiB in profiles/r01_v2b_*), matches whose
// source lies inside the batch read it back at shared-memory latency, and the finished tile goes
// to global memory with aligned 16-byte stores.
template <bool GATED>
__device__ __forceinline__ bool copy_batch(const uint8_t *__restrict__ sg, uint8_t *og, uint32_t opg,
					   uint32_t capg, const SeqDesc *sdg, uint32_t c, int lane,
					   uint8_t *tile, uint32_t &total, const BatchGate *gate = nullptr)
{
	const bool act = static_cast<uint32_t>(lane) < c;
	uint32_t lit_pos = 0, lit = 0, ml = 0;
	if (act) {
		const uint2 raw = *reinterpret_cast<const uint2 *>(sdg + lane);
		lit_pos = raw.x;
		lit = raw.y & 0xffffu;
		ml = raw.y >> 16;
	}
	uint32_t off = 0;
	if (ml) {
		const uint8_t *qp = sg + lit_pos + lit;
		off = ld_u8<true>(qp) | (ld_u8<true>(qp + 1) << 8);
	}
	const uint32_t len = lit + ml;
	uint32_t incl = len;
#pragma unroll
	for (int k = 1; k < 32; k <<= 1) {
		const uint32_t v = __shfl_up_sync(FULL_MASK, incl, k);
		if (lane >= k) incl += v;
	}
	total = __shfl_sync(FULL_MASK, incl, 31);
	const uint32_t out_pos = opg + incl - len;   // block-relative start of this sequence's output
	const uint32_t mo = out_pos + lit;           // ... and of its match
	const bool bad = ml && (off == 0 || off > mo);
	if (__any_sync(FULL_MASK, bad) || total > capg - opg) return false;

	const uint32_t src_s = mo - off;
	const uint32_t src_e = src_s + (ml < off ? ml : off);   // self-overlap: source ends where the match starts
	// tile mode needs every source to lie entirely before the batch or entirely inside it
	const bool awkward = ml && src_s < opg && (src_e > opg || off < ml);
	const bool use_tile = total <= TILE_BYTES && !__any_sync(FULL_MASK, awkward);
	// out(x): where output byte x (block-relative) is assembled.  The tile keeps the 16-byte phase
	// of the global address so that the flush is aligned on both sides.
	const uint32_t pad = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(og + opg) & 15);
	uint8_t *const wbase = use_tile ? tile + pad - opg : og;

	// ---- literals: lane per sequence (text-like data: mean ~2 bytes, rarely > 16) ----
	const uint32_t maxlit = __reduce_max_sync(FULL_MASK, lit);
	if (maxlit) {
		const uint8_t *lp = sg + lit_pos;
		uint8_t *dp = wbase + out_pos;
		const uint32_t lim = maxlit < 16 ? maxlit : 16;
		for (uint32_t k = 0; k < lim; k++)
			if (k < lit) dp[k] = static_cast<uint8_t>(ld_u8<true>(lp + k));
		if (maxlit > 16) {
			uint32_t big = __ballot_sync(FULL_MASK, lit > 16);
			while (big) {
				const int j = __ffs(big) - 1;
				big &= big - 1;
				warp_copy<true>(wbase + __shfl_sync(FULL_MASK, out_pos, j) + 16,
						sg + __shfl_sync(FULL_MASK, lit_pos, j) + 16,
						__shfl_sync(FULL_MASK, lit, j) - 16, lane);
			}
		}
	}

	// ---- matches ----
	// A match may start once every earlier sequence of this batch whose output overlaps its source
	// has finished (bytes before the batch are final, literals of the batch are written above).
	// dep = the lanes it waits for, found by binary search over the monotone output positions.
	bool done = (ml == 0);
	uint32_t dep = 0;
	if (__any_sync(FULL_MASK, !done && src_e > opg)) {
		// lo = number of sequences that end at or before src_s; hi = number that start before src_e
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
		if (done || src_e <= opg) dep = 0;
	}
	// where this lane's match source is read from: final global output, or the tile
	const uint8_t *const sp = (use_tile && src_s >= opg) ? tile + pad + (src_s - opg) : og + src_s;
	uint8_t *const dp = wbase + mo;
	const bool simple_kind = ml <= 32 && off >= ml;
	__syncwarp();
	// round 1: every short, non-overlapping match that waits for nothing inside the batch
	auto parallel_round = [&](bool simple) {
		if (__any_sync(FULL_MASK, simple)) {
			const uint32_t maxml = __reduce_max_sync(FULL_MASK, simple ? ml : 0u);
			if (simple) {
				const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(sp) & 15);
				const uint4 *base = reinterpret_cast<const uint4 *>(sp - m);
				const uint4 A = base[0];
				uint4 B = make_uint4(0, 0, 0, 0), C = make_uint4(0, 0, 0, 0);
				if (m + ml > 16) B = base[1];
				if (m + ml > 32) C = base[2];
				unsigned long long d0 = A.x | (static_cast<unsigned long long>(A.y) << 32);
				unsigned long long d1 = A.z | (static_cast<unsigned long long>(A.w) << 32);
				unsigned long long d2 = B.x | (static_cast<unsigned long long>(B.y) << 32);
				unsigned long long d3 = B.z | (static_cast<unsigned long long>(B.w) << 32);
				unsigned long long d4 = C.x | (static_cast<unsigned long long>(C.y) << 32);
				const unsigned long long d5 = C.z | (static_cast<unsigned long long>(C.w) << 32);
				if (m & 8) { d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = d5; }
				const uint32_t sh = (m & 7) * 8;
				unsigned long long r[4] = {d0, d1, d2, d3};
				if (sh) {
					r[0] = (d0 >> sh) | (d1 << (64 - sh));
					r[1] = (d1 >> sh) | (d2 << (64 - sh));
					r[2] = (d2 >> sh) | (d3 << (64 - sh));
					r[3] = (d3 >> sh) | (d4 << (64 - sh));
				}
#pragma unroll
				for (int c8 = 0; c8 < 4; c8++) {
					if (static_cast<uint32_t>(c8 * 8) < maxml) {
#pragma unroll
						for (int k = 0; k < 8; k++)
							if (static_cast<uint32_t>(c8 * 8 + k) < ml)
								dp[c8 * 8 + k] = static_cast<uint8_t>(r[c8] >> (8 * k));
					}
				}
			}
		}
		done = done || simple;
	};
	if (GATED) {
		// sources that end before the chain's final frontier may go now; the rest of this batch's
		// past becomes final when every earlier batch has finished
		uint32_t d = 0, fpos = 0;
		if (lane == 0) {
			// the ring slot of batch d is only reused once batch d is done: re-check the frontier
			for (;;) {
				d = *gate->done_upto;
				if (d >= gate->my_batch) {
					fpos = opg;
				} else {
					const uint32_t sl = d % gate->slots;
					// positions are frame-relative: a frontier inside an earlier frame says nothing about mine
					const bool same = gate->base_lo[sl] == gate->my_base_lo && gate->base_hi[sl] == gate->my_base_hi;
					fpos = same ? gate->out_start[sl] : 0u;
				}
				if (d >= gate->my_batch || *gate->done_upto == d) break;
			}
		}
		d = __shfl_sync(FULL_MASK, d, 0);
		fpos = __shfl_sync(FULL_MASK, fpos, 0);
		if (d < gate->my_batch) {
			parallel_round(!done && dep == 0 && simple_kind && src_e <= fpos);
			if (lane == 0) {
				while (*gate->done_upto < gate->my_batch) __nanosleep(40);
			}
			__syncwarp();
			__threadfence_block();
		}
	}
	parallel_round(!done && dep == 0 && simple_kind);
	// everything else strictly in sequence order, one match at a time by the whole warp: matches that
	// wait for output of this batch (their source is in the tile: shared-memory latency), long and
	// self-overlapping ones
	uint32_t rest = __ballot_sync(FULL_MASK, !done);
	if (rest) {
		// One packed word per lane so that a single shuffle serves the common case: a short,
		// non-overlapping match whose source and destination both live in the tile.
		const bool quick = use_tile && simple_kind && src_s >= opg;
		const uint32_t packed = quick ? (0x80000000u | ((mo - opg) << 17) | ((src_s - opg) << 6) | (ml - 1)) : 0u;
		uint8_t *const tp = tile + pad;
		while (rest) {
			const int j = __ffs(rest) - 1;
			rest &= rest - 1;
			const uint32_t pk = __shfl_sync(FULL_MASK, packed, j);
			__syncwarp();
			if (pk & 0x80000000u) {
				const uint32_t mlj = (pk & 63u) + 1u;
				if (static_cast<uint32_t>(lane) < mlj) tp[((pk >> 17) & 0x3fffu) + lane] = tp[((pk >> 6) & 0x7ffu) + lane];
				continue;
			}
			const uint32_t offj = __shfl_sync(FULL_MASK, off, j), mlj = __shfl_sync(FULL_MASK, ml, j);
			const uint32_t moj = __shfl_sync(FULL_MASK, mo, j);
			const uint32_t ssj = moj - offj;
			uint8_t *dj = wbase + moj;
			const uint8_t *sj = (use_tile && ssj >= opg) ? tile + pad + (ssj - opg) : og + ssj;
			if (offj >= mlj) {
				if (mlj <= 32) {
					if (static_cast<uint32_t>(lane) < mlj) dj[lane] = sj[lane];
				} else {
					warp_copy<false>(dj, sj, mlj, lane);
				}
			} else {
				match_copy_overlap(dj, offj, mlj, lane);   // source directly in front of dj
			}
		}
	}
	__syncwarp();
	if (use_tile) {
		// flush: tile[pad .. pad + total) -> og[opg ..], 16-byte aligned on both sides
		uint8_t *g = og + opg;
		const uint8_t *t = tile + pad;
		const uint32_t head = (16u - pad) & 15u;
		const uint32_t h = head < total ? head : total;
		if (static_cast<uint32_t>(lane) < h) g[lane] = t[lane];
		const uint32_t rest = total - h;
		const uint32_t nvec = rest >> 4;
		const uint4 *t4 = reinterpret_cast<const uint4 *>(t + h);
		uint4 *g4 = reinterpret_cast<uint4 *>(g + h);
		for (uint32_t v = lane; v < nvec; v += 32) g4[v] = t4[v];
		const uint32_t tail = rest & 15, tb = h + (nvec << 4);
		if (static_cast<uint32_t>(lane) < tail) g[tb + lane] = t[tb + lane];
		__syncwarp();
	}
	return true;
}

// One token of the fast path: ip = token position.  On success p = first literal byte, lit / ml the
// lengths (ml = 0 for the final literal-only sequence), nxt = next token position.
__device__ __forceinline__ bool parse_token(const uint8_t *__restrict__ s, uint32_t n, uint32_t ip, uint32_t &p,
					    uint32_t &lit, uint32_t &ml, uint32_t &nxt)
{
	const uint32_t t = ld_u8<true>(s + ip);
	lit = t >> 4;
	ml = t & 15;
	p = ip + 1;
	if (lit == 15 || ml == 15) return parse_extended(s, n, p, lit, ml, nxt);
	const uint32_t q = p + lit;
	if (q + 2 <= n) {
		ml += 4;
		nxt = q + 2;
		return true;
	}
	if (q == n && ml == 0) {
		nxt = n;
		return true;
	}
	return false;
}

// One token with lengths of any size, by the whole warp: the extension bytes are scanned 32 at a time
// (read_length_ext; a 4 MiB run has 16 448 of them).  Same results as parse_token, without its 16-bit limit on the
// fields; false for truncation and for lengths beyond 2^31.  Warp-uniform.
__device__ __forceinline__ bool parse_token_wide(const uint8_t *__restrict__ s, uint32_t n, uint32_t ip, uint32_t &p,
						 uint32_t &lit, uint32_t &ml, uint32_t &nxt, int lane)
{
	const uint32_t t = ld_u8<true>(s + ip);
	lit = t >> 4;
	ml = t & 15;
	p = ip + 1;
	if (lit == 15) {
		if (!read_length_ext(s, p, n, lit, lane) || lit > 0x7fffffffu) return false;
	}
	if (lit > n - p) return false;
	const uint32_t q = p + lit;
	if (q == n) {
		if (ml) return false;
		nxt = n;
		return true;
	}
	if (q + 2 > n) return false;
	nxt = q + 2;
	if (ml == 15) {
		if (!read_length_ext(s, nxt, n, ml, lane) || ml > 0x7ffffff0u) return false;
	}
	ml += 4;
	return true;
}

// One block of a chain by one warp through the batch machinery (lane 0 walks the token chain).
// Positions are relative to `frame_out` (the start of the frame's flat output), so a match may
// reach back across block boundaries; pos advances by what the block produced.  Returns false when
// the block needs the exact routine.
__device__ __forceinline__ bool chain_block_fast(const uint8_t *__restrict__ s, uint32_t n, uint8_t *frame_out,
						 uint32_t &pos, uint32_t cap_abs, SeqDesc *sd, uint8_t *tile,
						 int lane)
{
	uint32_t ip = 0;
	while (ip < n) {
		uint32_t cnt = 0;
		bool fb = false;
		if (lane == 0) {
			while (cnt < 32 && ip < n) {
				const uint32_t t = ld_u8<true>(s + ip);
				uint32_t lit = t >> 4, ml = t & 15, p = ip + 1, nxt;
				if (lit == 15 || ml == 15) {
					if (!parse_extended(s, n, p, lit, ml, nxt)) { fb = true; break; }
				} else {
					const uint32_t q = p + lit;
					if (q + 2 <= n) {
						ml += 4;
						nxt = q + 2;
					} else if (q == n && ml == 0) {
						nxt = n;
					} else {
						fb = true;
						break;
					}
				}
				*reinterpret_cast<uint2 *>(sd + cnt) = make_uint2(p, lit | (ml << 16));
				cnt++;
				ip = nxt;
			}
		}
		__syncwarp();
		fb = __shfl_sync(FULL_MASK, fb ? 1 : 0, 0) != 0;
		cnt = __shfl_sync(FULL_MASK, cnt, 0);
		ip = __shfl_sync(FULL_MASK, ip, 0);
		if (fb) return false;
		if (cnt == 0) break;
		uint32_t total = 0;
		if (!copy_batch<false>(s, frame_out, pos, cap_abs, sd, cnt, lane, tile, total)) return false;
		pos += total;
		__syncwarp();
	}
	return true;
}

// One warp, G blocks (first_block .. first_block + G - 1).  sd = this warp's [G][32] descriptors.
template <int G>
__device__ __forceinline__ void decode_group(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
					     uint32_t first_block, const lz4b200_blk_desc *__restrict__ desc,
					     lz4b200_blk_status *status, SeqDesc *sd, uint8_t *tile, int lane)
{
	static_assert(G >= 1 && G <= 32, "one lane per block");
	uint32_t state = LS_IDLE;
	const uint8_t *s = src;
	uint8_t *o = dst;
	uint32_t n = 0, cap = 0, flags = 0, ip = 0, op = 0;
	uint32_t computed = 0, declared = 0, code = LZ4B200_ST_OK;
	if (lane < G && first_block + lane < n_blocks) {
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

	// ---- fused block checksum: quad q hashes block 8 * pass + q, eight chains per warp at a time
	//      (lib/lz4ada.adb:698-707) ----
#pragma unroll 1
	for (int pass = 0; pass * 8 < G; pass++) {
		const int q = lane >> 2;
		const int blk = pass * 8 + q;
		const int qq = blk < G ? blk : 0;
		const uint8_t *sq = shfl_cptr(s, qq);
		const uint32_t nq = __shfl_sync(FULL_MASK, n, qq);
		const uint32_t fq = __shfl_sync(FULL_MASK, flags, qq);
		const uint32_t stq = __shfl_sync(FULL_MASK, state, qq);
		const bool want = blk < G && stq == LS_RUN && (fq & LZ4B200_BLK_HAS_CHECKSUM);
		if (__any_sync(FULL_MASK, want)) {
			const uint32_t h = quad_xxh32_prologue(sq, want ? nq : 0, lane);
			// lane g (block g) of this pass picks up the digest of its quad
			const int src_lane = ((lane - pass * 8) * 4) & 31;
			const uint32_t hq = __shfl_sync(FULL_MASK, h, src_lane);
			if (lane >= pass * 8 && lane < pass * 8 + 8 && state == LS_RUN && (flags & LZ4B200_BLK_HAS_CHECKSUM)) {
				const uint8_t *t = s + n;
				declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) |
					   (ld_u8<true>(t + 3) << 24);
				computed = hq;
				if (computed != declared) {
					code = LZ4B200_ST_BLOCK_CHECKSUM;
					state = LS_DONE;
				}
			}
		}
	}
	if (state == LS_RUN && (flags & LZ4B200_BLK_HASH_ONLY)) state = LS_DONE;

	// ---- stored blocks: warp-cooperative copy (lib/lz4ada.adb:685-695) ----
#pragma unroll 1
	for (int g = 0; g < G; g++) {
		const uint32_t stg = __shfl_sync(FULL_MASK, state, g);
		const uint32_t fg = __shfl_sync(FULL_MASK, flags, g);
		if (stg != LS_RUN || !(fg & LZ4B200_BLK_STORED)) continue;
		const uint32_t ng = __shfl_sync(FULL_MASK, n, g);
		const uint32_t capg = __shfl_sync(FULL_MASK, cap, g);
		if (ng > capg) {
			if (lane == g) state = LS_FALLBACK;
			continue;
		}
		warp_copy<true>(shfl_ptr(o, g), shfl_cptr(s, g), ng, lane);
		if (lane == g) {
			op = ng;
			state = LS_DONE;
		}
	}

	// ---- compressed blocks: alternate parse (lane per block) and copy (lane per sequence) ----
	for (;;) {
		if (state == LS_RUN && ip >= n) state = LS_DONE;
		if (!__any_sync(FULL_MASK, state == LS_RUN)) break;
		uint32_t cnt = 0;
		if (state == LS_RUN) {
			SeqDesc *my = sd + lane * SD_STRIDE;
			bool fb = false;
			while (cnt < 32 && ip < n) {
				const uint32_t t = ld_u8<true>(s + ip);
				uint32_t lit = t >> 4, ml = t & 15, p = ip + 1, nxt;
				if (lit == 15 || ml == 15) {
					// length extensions: out of line, rare on text-like data
					if (!parse_extended(s, n, p, lit, ml, nxt)) { fb = true; break; }
				} else {
					const uint32_t q = p + lit;
					if (q + 2 <= n) {          // offset present (the reference needs both bytes, :766)
						ml += 4;
						nxt = q + 2;
					} else if (q == n && ml == 0) {   // final literal-only sequence
						nxt = n;
					} else {
						fb = true;
						break;
					}
				}
				*reinterpret_cast<uint2 *>(my + cnt) = make_uint2(p, lit | (ml << 16));
				cnt++;
				ip = nxt;
			}
			if (fb) {
				state = LS_FALLBACK;
				cnt = 0;
			}
		}
		__syncwarp();
#pragma unroll 1
		for (int g = 0; g < G; g++) {
			const uint32_t c = __shfl_sync(FULL_MASK, cnt, g);
			if (c == 0) continue;
			uint32_t total = 0;
			const bool okay = copy_batch<false>(shfl_cptr(s, g), shfl_ptr(o, g), __shfl_sync(FULL_MASK, op, g),
						     __shfl_sync(FULL_MASK, cap, g), sd + g * SD_STRIDE, c, lane, tile, total);
			if (lane == g) {
				if (okay) op += total;
				else state = LS_FALLBACK;
			}
		}
		__syncwarp();
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
	for (int g = 0; g < G; g++) {
		if (__shfl_sync(FULL_MASK, state, g) != LS_FALLBACK) continue;
		const uint32_t b = first_block + g;
		const lz4b200_blk_desc d = desc[b];
		process_block<false>(src, dst + d.dst_off, d, d.dst_cap, d.hist_avail, status + b, lane);
	}
}

}  // namespace lz4b200
