// kernels_v6.cuh -- K1 sixth generation: one LANE per block like v5, rebuilt around what v5's profile said.
//
// v5 (profiles/r01_s2_k1_ncu_full_4gib.csv): 9.8 G warp instructions = ~760 per trip of the lock-step loop, because
// one trip carried every stage of a sequence at its worst-case size (16 literal bytes, 32 match bytes, a match piece
// deferred across the literals of the next sequence), and a trip could not be shorter than the DRAM round trip of
// the old match source it had asked for one trip earlier.  v6 changes the decomposition of the *work*, not of the
// data:
//
//   piece     the unit of a trip is one piece of a sequence: <= 8 literal bytes followed by <= 16 match bytes
//             (min(16, distance): a piece never overlaps its source; for periods below 16 the distance doubles from
//             piece to piece -- pattern replication, lib/lz4ada.adb:893-903).  A sequence with <= 7 literals and a
//             match of <= 16 bytes (90 % of them on text) is ONE trip; longer runs take more trips of that lane
//   parse     runs K pieces AHEAD of the copy: token, lengths and offset need the compressed bytes only
//             (Decompress_Sequence, :737-777).  The parse side pushes a piece descriptor into a register FIFO that
//             every lane shifts once per trip (warp-uniform depth, no per-lane head / tail), literal bytes travel
//             inside the descriptor, and an old match source is requested with cp.async into the staging slot of
//             that trip -- K trips before the copy side needs it, so no trip waits for memory
//   copy      pops the oldest descriptor: bytes from the descriptor (literals), from the lane's out ring (young
//             source) or from the staging slot (old source) -> funnel shifts -> the out ring at the cursor, strictly
//             in output order (no holes, nothing to keep behind a store); one complete 16-byte chunk per trip goes to
//             global memory
//   in ring   128 bytes per lane, row layout (a lane's ring is contiguous, 144-byte stride: one 16-byte cp.async per
//             refill; the first chunk is mirrored behind the last so that reads never wrap)
//   out ring  the lane's last 256 / 512 bytes, word w of lane L at word w * 32 + L (bank = lane: conflict-free)
//
// Block checksums (Check_Checksum, :698-707) are folded in per lane as the chunks arrive, as in v5.  Anything the
// fast path does not take -- stored blocks, length fields longer than the in ring shows, every error -- goes to
// process_block (exact semantics of :716-904) for the whole block.
#pragma once

#include "kernels_v2.cuh"

#ifndef LZ4B200_V6_HINTS
#define LZ4B200_V6_HINTS 3   // bit 0: old match sources evict_first, bit 1: output stores evict_last, bit 2: compressed input evict_first
#endif

namespace lz4b200 {
namespace v6 {

constexpr uint32_t IN_BYTES = 128;         // in ring per lane
constexpr uint32_t IN_STRIDE = 144;        // + one mirror chunk
constexpr uint32_t LIT_PIECE = 8, ML_PIECE = 16;
constexpr uint32_t VIEW = 12;              // bytes of the compressed stream a trip looks at
constexpr uint32_t BACKLOG = 96;           // parse waits while more than this is parsed but not yet flushed
constexpr uint32_t STAGE_SLOT = 32;        // bytes per lane and staging slot: the two 16-byte granules around a source

enum : uint32_t { KIND_NEAR = 0, KIND_FAR = 1 };   // where the match bytes of a piece come from
enum : uint32_t { L_IDLE = 0, L_RUN = 1, L_EXACT = 2 };

template <uint32_t OWW, int K> struct Layout {
	static constexpr uint32_t OUT_BYTES = OWW * 4;                      // out ring per lane
	static constexpr uint32_t OUT_WARP = OWW * 128;                     // ... per warp; rings are aligned to this
	static constexpr uint32_t NEAR_MAX = OUT_BYTES - ML_PIECE - 8;      // young sources: distance <= this (> BACKLOG + 40)
	static constexpr uint32_t IN_WARP = 32 * IN_STRIDE;
	static constexpr uint32_t STAGE_WARP = K * 32 * STAGE_SLOT;
	static constexpr uint32_t SIDE_WARP = IN_WARP + STAGE_WARP;         // in rings + staging, per warp
	// dynamic shared memory of a CTA of `warps`: side areas first, then the out rings at the next multiple of
	// OUT_WARP (the address of ring word j is then base | ((u + 128 j) & mask): one LOP3)
	static constexpr uint32_t smem_bytes(uint32_t warps) { return warps * SIDE_WARP + OUT_WARP + warps * OUT_WARP; }
};

__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
template <int OFF> __device__ __forceinline__ uint32_t lds32o(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
	return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cp16(uint32_t sa, const void *g)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
}
// ... with an L2 eviction policy (createpolicy): old match sources are read once and must not push the recently
// written output out of the L2, which is where the next matches look for it
__device__ __forceinline__ void cp16_hint(uint32_t sa, const void *g, uint64_t pol)
{
	asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg128_hint(void *g, const uint4 &v, uint64_t pol)
{
	asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(g), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// One warp: lanes take blocks from *counter until it reaches n_blocks.
template <uint32_t OWW, int K>
__device__ __forceinline__ void decode_lanes(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
					     const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, uint32_t *counter,
					     uint32_t in_base /* shared address of this lane's in ring */,
					     uint32_t stage_base /* ... of this lane's 32 bytes in staging slot 0 */,
					     uint32_t ring_base /* ... of this warp's out rings (aligned to their size) + lane * 4 */, int lane)
{
	// L2 eviction hints, measured on the headline corpus (tools/probes/v6_ncu_hints.sh): old match sources marked
	// evict_first + output stores evict_last = 28.7 GB of DRAM reads per launch instead of 33.9 GB, hit rate 21.5 -> 27 %
	constexpr uint32_t hints = LZ4B200_V6_HINTS;
	using L = Layout<OWW, K>;
	uint64_t pol_first, pol_last;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
	constexpr uint32_t M = L::OUT_WARP - 1;
	const uint32_t RB = ring_base & ~M;   // ring word j of this lane: RB | ((u + 128 j) & M), u = ((x >> 2) << 7) + ring_base

	// ---- lane state.  Input positions are "aligned stream coordinates": position of a payload byte = its offset in
	// the block + mis, so that position 16 c starts the c-th aligned 16-byte chunk in global memory (sbase[x] is the
	// byte at position x).  Output positions likewise: offset in the block's output + the misalignment of its start.
	uint32_t state = L_IDLE, blk = 0;
	const uint8_t *sbase = src;
	uint8_t *obase = dst;
	// parse side: the sequence in progress is (rem_l literal bytes, then -- need_off -- its offset, then rem_m match
	// bytes at distance dist); all three zero = the cursor stands on a token
	uint32_t a = 0, a_end = 0, a_req = 0, ifl = 0;   // cursor, end, next chunk to request, requests of the last trips (bit t)
	uint32_t q = 0, p_start = 0, p_cap = 0;          // output position behind every piece pushed so far
	uint32_t rem_l = 0, rem_m = 0, need_off = 0, dist = 0, mln = 0, drain = 0;
	bool ended = false;                              // the block is parsed to its end: drain pieces in flight are counted down
	// copy side
	uint32_t p = 0, p_fl = 0;                        // output cursor, flush frontier (multiple of 16)
	// fused block checksum (XXHash32.Process, lib/lz4ada.adb:979-991)
	uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, h_pos = 0, a_beg = 0, declared = 0;
	bool hashing = false;
	// descriptor FIFO: [0] is popped by the copy side, [K - 1] was pushed last.
	// descriptor = kind | literal bytes << 2 | match bytes << 6 | (distance or source misalignment) << 16
	uint32_t f_desc[K], f_d0[K], f_d1[K];
#pragma unroll
	for (int i = 0; i < K; i++) f_desc[i] = f_d0[i] = f_d1[i] = 0;
	bool exhausted = false;
	uint32_t trip = 0, slot_i = 0, idle_trips = 0;

	for (;;) {
		// ================= lanes that are not decoding: new blocks, the exact routine, the end =================
		const uint32_t special = __ballot_sync(FULL_MASK, state != L_RUN);
		if (special) {
			// new blocks when enough lanes are idle (half a warp at a time: what is left of a batch ends up in few,
			// full warps instead of a few lanes of each)
			const uint32_t idle = __ballot_sync(FULL_MASK, state == L_IDLE);
			if (idle && !exhausted && (__popc(idle) >= 16 || special == FULL_MASK)) {
				cp_async_wait<0>();   // nothing of a lane's previous block may still be landing in its rings
				uint32_t base = 0;
				const uint32_t want = __popc(idle);
				if (lane == 0) base = atomicAdd(counter, want);
				base = __shfl_sync(FULL_MASK, base, 0);
				if (base + want >= n_blocks) exhausted = true;
				if (state == L_IDLE) {
					const uint32_t b = base + __popc(idle & ((1u << lane) - 1u));
					if (b < n_blocks) {
						const lz4b200_blk_desc d = desc[b];
						if (!(d.flags & LZ4B200_BLK_NOT_K1)) {
							blk = b;
							const uint8_t *hs = src + d.src_off;
							const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(hs) & 15u);
							uint8_t *og = dst + d.dst_off;
							const uint32_t oph = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(og) & 15u);
							sbase = hs - mis;
							obase = og - oph;
							a = a_beg = h_pos = mis;
							a_end = mis + d.src_len;
							a_req = 0;
							ifl = 0;
							q = p = p_start = oph;
							p_cap = oph + d.dst_cap;
							p_fl = 0;   // chunk 0 is partial when oph != 0: flushed bytewise
							rem_l = rem_m = need_off = 0;
							ended = false;
							acc0 = PRIME_1 + PRIME_2; acc1 = PRIME_2; acc2 = 0; acc3 = 0u - PRIME_1;   // Reset, :932-940
							// positions are 32-bit with headroom; stored blocks and the like go to the exact routine
							const bool plain = !(d.flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_HASH_ONLY)) && d.dst_cap < 0x7fff0000u &&
									   d.src_len < 0x7fff0000u;
							state = plain ? L_RUN : L_EXACT;
							hashing = plain && (d.flags & LZ4B200_BLK_HAS_CHECKSUM);
							if (hashing) {
								const uint8_t *t = hs + d.src_len;
								declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
							}
						}
					}
				}
				__syncwarp();
			}
			if (__all_sync(FULL_MASK, state == L_IDLE)) {
				if (exhausted) break;
				continue;
			}
			// blocks for the exact routine (whole warp)
			uint32_t coop = __ballot_sync(FULL_MASK, state == L_EXACT);
			if (coop) {
				cp_async_wait<0>();
				ifl = 0;
				while (coop) {
					const int j = __ffs(coop) - 1;
					coop &= coop - 1;
					const uint32_t bj = __shfl_sync(FULL_MASK, blk, j);
					const lz4b200_blk_desc d = desc[bj];
					process_block<false>(src, dst + d.dst_off, d, d.dst_cap, d.hist_avail, status + bj, lane);
					if (lane == j) state = L_IDLE;
					__syncwarp();
				}
				continue;
			}
		}

		const uint32_t slot = stage_base + slot_i * (32 * STAGE_SLOT);   // staging slot of this trip (warp-uniform)
		const bool run = state == L_RUN;
		bool progressed = false, bad = false;

		cp_async_wait<K - 1>();   // the staging slot of the piece pushed K trips ago, and the in-ring chunk requested with it, have landed
		// the twelve bytes at the parse cursor, requested before the copy side's work so that the parse arithmetic can be
		// scheduled under it (shared-memory accesses keep their program order; arithmetic does not)
		const uint32_t va = in_base + (a & (IN_BYTES - 4u));
		const uint32_t w0 = lds32o<0>(va), w1 = lds32o<4>(va), w2 = lds32o<8>(va), w3 = lds32o<12>(va);
		// ================= copy side: the piece pushed K trips ago =================
		{
			const uint32_t dsc = f_desc[0];
			const uint32_t nl = (dsc >> 2) & 15u, n = (dsc >> 6) & 31u, info = dsc >> 16;
			// ---- literals, out of the descriptor (Write_Output, :790-824) ----
			if (run && nl) {
				const uint32_t d0 = f_d0[0], d1 = f_d1[0];
				const uint32_t k8 = (p & 3u) * 8u;
				const uint32_t u = ((p >> 2) << 7) + ring_base;
				const uint32_t a0 = (u & M) | RB;
				const uint32_t nw = ((p & 3u) + nl + 3u) >> 2;   // words touched, 1 .. 3
				const uint32_t old0 = lds32(a0);
				sts32(a0, (d0 << k8) | (old0 & ((1u << k8) - 1u)));
				if (nw > 1) sts32(((u + 128u) & M) | RB, __funnelshift_l(d0, d1, k8));
				if (nw > 2) sts32(((u + 256u) & M) | RB, __funnelshift_l(d1, 0u, k8));
				p += nl;
			}
			// ---- match bytes (Output_With_History, :845-904): young source = the out ring, old = the staging slot ----
			if (run && n) {
				const bool far = (dsc & 3u) == KIND_FAR;
				const uint32_t s = p - info;
				const uint32_t us = ((s >> 2) << 7) + ring_base;
				const uint32_t sa = slot + (info & 12u);
				const uint32_t bs = ((far ? info : s) & 3u) * 8u;
				const uint32_t x0 = lds32(far ? sa : ((us & M) | RB));
				const uint32_t x1 = lds32(far ? sa + 4u : (((us + 128u) & M) | RB));
				const uint32_t x2 = lds32(far ? sa + 8u : (((us + 256u) & M) | RB));
				const uint32_t x3 = lds32(far ? sa + 12u : (((us + 384u) & M) | RB));
				const uint32_t x4 = lds32(far ? sa + 16u : (((us + 512u) & M) | RB));
				const uint32_t D0 = __funnelshift_r(x0, x1, bs), D1 = __funnelshift_r(x1, x2, bs);
				const uint32_t D2 = __funnelshift_r(x2, x3, bs), D3 = __funnelshift_r(x3, x4, bs);
				const uint32_t k8 = (p & 3u) * 8u;
				const uint32_t u = ((p >> 2) << 7) + ring_base;
				const uint32_t a0 = (u & M) | RB;
				const uint32_t nw = ((p & 3u) + n + 3u) >> 2;   // words touched, 1 .. 5
				const uint32_t old0 = lds32(a0);
				sts32(a0, (D0 << k8) | (old0 & ((1u << k8) - 1u)));
				if (nw > 1) sts32(((u + 128u) & M) | RB, __funnelshift_l(D0, D1, k8));
				if (nw > 2) sts32(((u + 256u) & M) | RB, __funnelshift_l(D1, D2, k8));
				if (nw > 3) sts32(((u + 384u) & M) | RB, __funnelshift_l(D2, D3, k8));
				if (nw > 4) sts32(((u + 512u) & M) | RB, __funnelshift_l(D3, 0u, k8));
				p += n;
			}
			progressed = run && (nl | n) != 0;
		}
		// shift the FIFO (the new piece goes to [K - 1] below)
#pragma unroll
		for (int i = 0; i + 1 < K; i++) {
			f_desc[i] = f_desc[i + 1];
			f_d0[i] = f_d0[i + 1];
			f_d1[i] = f_d1[i + 1];
		}
		uint32_t n_desc = 0, n_d0 = 0, n_d1 = 0;

		// ================= parse side: one piece (Decompress_Sequence, lib/lz4ada.adb:737-777) =================
		{
			const uint32_t a_ld = a_req - 16u * __popc(ifl & ((1u << (K - 1)) - 1u));   // chunks below this have landed
			bool can = run && !ended && (a_ld >= a_end || a_ld >= a + VIEW) && q - p_fl <= BACKLOG;
			const uint32_t sh = (a & 3u) * 8u;
			const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
			const bool fresh = (rem_l | rem_m | need_off) == 0;   // the cursor stands on a token
			// ---- rare: a length whose first extension byte is 255 (runs of 270 literals / matches of 274 bytes and more,
			// Decompress_Sequence's length loops, :741-747 and :773-777).  The bytes behind the first are summed here, out of
			// the in ring, in a loop of their own; the trip then consumes the whole length field and no payload.  A match
			// length is only taken this way by a trip that starts at the offset (a trip that finds one behind its literals
			// stops in front of the offset).  Fields longer than the ring can show go to the exact routine.
			const bool long_l = fresh && a + 1 < a_end && (v0 & 0xfff0u) == 0xfff0u;
			const bool long_m = need_off != 0 && rem_l == 0 && mln == 15u && a + 2 < a_end && ((v0 >> 16) & 255u) == 255u;
			uint32_t x_sum = 0, x_cnt = 0;
			bool x_over = false;
			if (can && (long_l || long_m)) {
				uint32_t pos = a + (long_l ? 2u : 3u);
				for (;;) {
					if (pos >= a_end || pos - a > 64u) { x_over = true; break; }
					if (pos >= a_ld) { can = false; break; }   // not landed yet: this lane sits the trip out
					const uint32_t b = (lds32(in_base + (pos & (IN_BYTES - 4u))) >> ((pos & 3u) * 8u)) & 255u;
					x_sum += b;
					x_cnt++;
					pos++;
					if (b != 255u) break;
				}
			}
			if (can) {
				// ---- token (only when no sequence is in progress) ----
				const bool end0 = fresh && a >= a_end;   // the block ends behind a match, or is empty
				const bool tok = fresh && !end0;
				const uint32_t tk = v0 & 255u, e1 = (v0 >> 8) & 255u;
				const bool ext_l = tok && tk >= 0xf0u;
				uint32_t o = tok ? (ext_l ? 2u + x_cnt : 1u) : 0u;   // bytes of the view consumed
				if (tok) {
					rem_l = (tk >> 4) + (ext_l ? e1 + x_sum : 0u);
					mln = tk & 15u;
					need_off = 1;
				}
				uint32_t why = 0;   // (statistics only)
				if ((ext_l && a + 1 >= a_end) || (long_l && x_over)) why |= 1u;
				if (tok && rem_l > a_end - a - o) why |= 2u;
				bad = why != 0;
				// ---- literals of this piece: they travel in the descriptor ----
				const uint32_t lim = long_l ? 0u : tok ? 7u : LIT_PIECE;
				const uint32_t nl = rem_l < lim ? rem_l : lim;
				n_d0 = __funnelshift_r(v0, v1, o * 8u);
				n_d1 = __funnelshift_r(v1, v2, o * 8u);
				rem_l -= nl;
				o += nl;
				// ---- offset and match length, once the literals are through ----
				const bool do_off = need_off != 0 && rem_l == 0 && !end0;
				const uint32_t ao = a + o;
				const bool fin_lit = do_off && ao >= a_end;   // final literal-only sequence (:752-764)
				const uint32_t wi = o >> 2;   // 32 bits at view byte o (o <= 9)
				const uint32_t xa = wi == 0 ? v0 : wi == 1 ? v1 : v2, xb = wi == 0 ? v1 : wi == 1 ? v2 : 0u;
				const uint32_t X = __funnelshift_r(xa, xb, (o & 3u) * 8u);
				const uint32_t off = X & 0xffffu, e2 = (X >> 16) & 255u;
				const bool ext_m = mln == 15u;
				const bool defer = ext_m && e2 == 255u && !long_m && ao + 2 < a_end;   // the next trip starts at the offset
				const bool has_m = do_off && !fin_lit && !defer;
				if (has_m) {
					rem_m = mln + 4u + (ext_m ? e2 + x_sum : 0u);
					dist = off;
					o += ext_m ? 3u + x_cnt : 2u;
					need_off = 0;
				}
				if (fin_lit && mln != 0) why |= 4u;
				if (has_m && (ao + 2 > a_end || off == 0)) why |= 8u;                       // :766-772
				if (has_m && off > q + nl - p_start) why |= 16u;                            // :864-874
				if (has_m && ext_m && (ao + 2 >= a_end || (long_m && x_over))) why |= 32u;
				bad = why != 0;
				if (end0 || fin_lit) {
					ended = true;
					need_off = 0;
					drain = K + 1;
				}
				a += o;
				// ---- a piece of the match: never overlaps its source ----
				uint32_t n = rem_m < ML_PIECE ? rem_m : ML_PIECE;
				n = n < dist ? n : dist;
				uint32_t kind = KIND_NEAR, info = dist;
				if (n != 0 && dist > L::NEAR_MAX && !bad) {
					// old source, flushed long ago (distance > NEAR_MAX > BACKLOG + 40): ask for the granules around it
					const uint8_t *g = obase + (q + nl - dist);
					const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g) & 15u);
					if (hints & 1u) {
						cp16_hint(slot, g - m, pol_first);
						if (m + n > 16u) cp16_hint(slot + 16u, g - m + 16, pol_first);
					} else {
						cp16(slot, g - m);
						if (m + n > 16u) cp16(slot + 16u, g - m + 16);
					}
					kind = KIND_FAR;
					info = m;
				}
				rem_m -= n;
				if (dist < ML_PIECE && n == dist) dist <<= 1;   // the pattern has doubled
				if (nl + n > p_cap - q) { why |= 64u; bad = true; }   // the exact routine reports the overflow
				if (bad) {
					counter[3] = blk;
					counter[4] = why;
					counter[5] = a - a_beg;
				}
				q += nl + n;
				n_desc = kind | (nl << 2) | (n << 6) | (info << 16);
				progressed = true;
			}
		}
		if (bad) {
			// hand the block to the exact routine (it decodes from the start and reports); forget what is queued
			atomicAdd(counter + 1, 1u);   // (statistics: lz4b200_k1_fallbacks)
			state = L_EXACT;
			n_desc = 0;
#pragma unroll
			for (int i = 0; i + 1 < K; i++) f_desc[i] = 0;
		}
		f_desc[K - 1] = n_desc;
		f_d0[K - 1] = n_d0;
		f_d1[K - 1] = n_d1;

		ifl <<= 1;
		if (trip & 1u) {
			// ================= odd trips: up to two complete 16-byte chunks to global memory ... =================
#pragma unroll
			for (int rep = 0; rep < 2; rep++) {
				if (state == L_RUN && p - p_fl >= 16u) {
					const uint32_t fa = ((((p_fl >> 2) << 7) + ring_base) & M) | RB;
					uint4 v;
					v.x = lds32o<0>(fa);
					v.y = lds32o<128>(fa);
					v.z = lds32o<256>(fa);
					v.w = lds32o<384>(fa);
					if (p_fl >= p_start) {
						if (hints & 2u) stg128_hint(obase + p_fl, v, pol_last);
						else *reinterpret_cast<uint4 *>(obase + p_fl) = v;
					} else {
						// the block's output starts in the middle of a 16-byte granule: bytes only
						const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll 1
						for (int k = 0; k < 16; k++)
							if (p_fl + k >= p_start) obase[p_fl + k] = static_cast<uint8_t>(ws[k >> 2] >> (8 * (k & 3)));
					}
					p_fl += 16;
					progressed = true;
				}
			}
			// ================= ... and one stripe of the fused block checksum, straight out of the in ring =================
			const uint32_t a_ld = a_req - 16u * __popc(ifl & ((1u << K) - 1u));
			if (state == L_RUN && hashing && h_pos + 16u <= a_end && h_pos + 16u <= a_ld) {
				const uint32_t va = in_base + (h_pos & (IN_BYTES - 4u));
				const uint32_t sh = (h_pos & 3u) * 8u;
				const uint32_t w0 = lds32o<0>(va), w1 = lds32o<4>(va), w2 = lds32o<8>(va), w3 = lds32o<12>(va), w4 = lds32o<16>(va);
				acc0 = xxh_round(acc0, __funnelshift_r(w0, w1, sh));
				acc1 = xxh_round(acc1, __funnelshift_r(w1, w2, sh));
				acc2 = xxh_round(acc2, __funnelshift_r(w2, w3, sh));
				acc3 = xxh_round(acc3, __funnelshift_r(w3, w4, sh));
				h_pos += 16;
				progressed = true;
			}
		} else {
			// ================= even trips: the in ring takes one aligned chunk while there is room =================
			const uint32_t a_end16 = (a_end + 15u) & ~15u;
			const uint32_t tail = (hashing && h_pos < a ? h_pos : a) & ~15u;   // the oldest byte still needed
			if (state == L_RUN && a_req < a_end16 && a_req + 16u - tail <= IN_BYTES) {
				const uint8_t *g = sbase + a_req;
				const uint32_t o = a_req & (IN_BYTES - 1u);
				if (hints & 4u) {
					cp16_hint(in_base + o, g, pol_first);
					if (o == 0) cp16_hint(in_base + IN_BYTES, g, pol_first);
				} else {
					cp16(in_base + o, g);
					if (o == 0) cp16(in_base + IN_BYTES, g);   // mirror of the first chunk behind the last
				}
				a_req += 16;
				ifl |= 1u;
			}
		}
		cp_async_commit();
		// ================= a block whose last piece has been copied: write what is left, report =================
		{
			bool fin = false;
			if (state == L_RUN && ended) {
				drain--;
				fin = drain == 0;
				progressed = true;
			}
			if (__any_sync(FULL_MASK, fin)) {
				if (fin) {
					// p == q: everything pushed has been copied.  Whole chunks, then the partial one bytewise.
					while (p_fl < p) {
						const uint32_t fa = ((((p_fl >> 2) << 7) + ring_base) & M) | RB;
						const uint32_t ws[4] = {lds32o<0>(fa), lds32o<128>(fa), lds32o<256>(fa), lds32o<384>(fa)};
						if (p_fl >= p_start && p_fl + 16u <= p) {
							*reinterpret_cast<uint4 *>(obase + p_fl) = make_uint4(ws[0], ws[1], ws[2], ws[3]);
						} else {
#pragma unroll 1
							for (int k = 0; k < 16; k++)
								if (p_fl + k >= p_start && p_fl + k < p) obase[p_fl + k] = static_cast<uint8_t>(ws[k >> 2] >> (8 * (k & 3)));
						}
						p_fl += 16;
					}
					uint32_t computed = 0;
					bool okay = true;
					if (hashing) {
						// the stripes the hash is still behind (their chunks are kept in the in ring), then Final (:993-1017)
						while (h_pos + 16u <= a_end) {
							const uint32_t va = in_base + (h_pos & (IN_BYTES - 4u));
							const uint32_t sh = (h_pos & 3u) * 8u;
							const uint32_t w0 = lds32o<0>(va), w1 = lds32o<4>(va), w2 = lds32o<8>(va), w3 = lds32o<12>(va), w4 = lds32o<16>(va);
							acc0 = xxh_round(acc0, __funnelshift_r(w0, w1, sh));
							acc1 = xxh_round(acc1, __funnelshift_r(w1, w2, sh));
							acc2 = xxh_round(acc2, __funnelshift_r(w2, w3, sh));
							acc3 = xxh_round(acc3, __funnelshift_r(w3, w4, sh));
							h_pos += 16;
						}
						computed = xxh_finish<true>(acc0, acc1, acc2, acc3, a_end - a_beg, sbase + h_pos, a_end - h_pos);
						okay = computed == declared;
					}
					lz4b200_blk_status *st = status + blk;
					st->code = okay ? LZ4B200_ST_OK : LZ4B200_ST_BLOCK_CHECKSUM;   // :672-676, :702
					st->out_len = okay ? p - p_start : 0u;
					st->err_pos = 0;
					st->aux = 0;
					st->xxh32_computed = computed;
					st->xxh32_declared = hashing ? declared : 0u;
					state = L_IDLE;
				}
			}
		}
		// safety net: a state the lock-step machine cannot leave must not hang the device -- hand the blocks to the
		// exact routine (a lane makes progress whenever it parses, copies, flushes, hashes, waits for a chunk or finishes)
		{
			const bool moving = progressed || state != L_RUN || (ifl & ((1u << K) - 1u)) != 0;
			idle_trips = __any_sync(FULL_MASK, moving) ? 0u : idle_trips + 1u;
			if (idle_trips > 4096u) {
				if (state == L_RUN) {
					atomicAdd(counter + 2, 1u);
					state = L_EXACT;
				}
				idle_trips = 0;
			}
		}
		trip++;
		slot_i = slot_i + 1 == static_cast<uint32_t>(K) ? 0u : slot_i + 1;
	}
	cp_async_wait<0>();
}

}  // namespace v6
}  // namespace lz4b200
