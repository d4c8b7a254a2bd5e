// kernels_v5.cuh -- K1 fifth generation: one LANE per block, 32 blocks in lock-step per warp.
//
// v2 / v4 give a block to a warp and a sequence to a lane.  ncu (profiles/r01_v4_*): 37 warp
// instructions per LZ4 sequence, 15-20 of 32 lanes active, because a sequence is ~10 bytes of work and
// everything around it (scans, dependency search, the serial hand-over of dependent matches) is
// per-batch overhead.  v5 turns the decomposition around, the way a CPU decodes: every lane walks its
// own block through a small state machine, one step per trip of a warp-uniform loop, so nothing is ever
// exchanged between lanes:
//
//   in ring   128 bytes of the lane's compressed stream, refilled with one aligned 16-byte global load
//             per trip when there is room
//   out ring  the lane's last 256 output bytes.  Literals and matches are assembled here; complete
//             16-byte chunks go to global memory with one aligned 16-byte store each (full sectors)
//   layout    word w of lane L lives at shared-memory word w * 32 + L: lane L only ever touches bank L,
//             so 32 lanes reading 32 unrelated positions is one conflict-free wavefront
//   a trip    [token] -> [<= 16 literal bytes] -> [offset] -> [<= 32 match bytes]; a typical sequence
//             passes all four stages in one trip, a long one takes more trips of that lane only.
//             A match is copied in pieces of min(32, distance) bytes; for periods below 32 the
//             distance doubles from piece to piece (pattern replication, lib/lz4ada.adb:893-903)
//   matches   source younger than ~200 bytes: out ring; older: three aligned 16-byte global loads
//             (flushed long ago); both land in registers, then one common routine stores <= 32 bytes
//   giants    literal runs / matches of 512 bytes and more: the lane parks and the whole warp does
//             that one copy in global memory with the v1 warp copies, then the lane re-seeds its rings
//
// Lanes fetch new blocks from a global counter when enough of them are idle, so unequal blocks do not
// leave lanes parked.  Block checksums (Check_Checksum, lib/lz4ada.adb:698-707) are fused into the pass over
// the compressed bytes: every lane folds each 16-byte stripe of its payload into its own XXH32 state straight
// out of the in ring, one stripe per trip, as the chunks arrive -- the payload is read from memory once.
// (A block with a giant skips parts of the in ring; its checksum is redone by a quad of the warp at the end.)
// Anything unusual goes to process_block (exact error semantics of lib/lz4ada.adb:716-904).
#pragma once

#include "kernels_v2.cuh"

namespace lz4b200 {
namespace v5 {

constexpr int WARPS = 2;                  // per CTA: small CTAs spread a batch that does not fill the chip evenly
constexpr uint32_t IWW = 32;              // in ring: words per lane (128 bytes)
constexpr uint32_t OWW = 64;              // out ring: words per lane (256 bytes)
constexpr uint32_t IN_BYTES = IWW * 4, OUT_BYTES = OWW * 4;
constexpr uint32_t LIT_PIECE = 16;        // literal bytes per trip
constexpr uint32_t ML_PIECE = 32;         // match bytes per trip
constexpr uint32_t GIANT = 512;           // runs of this length and more are done by the whole warp

struct __align__(16) WarpMem {
	uint32_t in_w[IWW][32];
	uint32_t out_w[OWW][32];
	uint4 stage[32][3];   // per lane: the 48 bytes around an old match source, landed by cp.async
};

// block level
enum : uint32_t { L_IDLE = 0, L_RUN = 1, L_GIANT = 2, L_EXACT = 3, L_FINISH = 4, L_HASH = 5 };
// block checksum of the lane's block
enum : uint32_t { H_NONE = 0, H_INLINE = 1, H_REDO = 2 };
// sequence level (while L_RUN)
enum : uint32_t { S_TOKEN = 0, S_LIT = 1, S_OFF = 2, S_MATCH = 3 };

struct Bytes36 { uint32_t w[9]; };   // up to 32 payload bytes starting at byte 0 (+ one word of slack for shifting)

// Column addressing: byte x of a lane's ring of W words is at ((x >> 2) mod W) * 128 + lane * 4 + (x & 3).
// u = (word << 7) | lane4 steps through the words by adding 128 and masking with W * 128 - 1.
template <uint32_t W>
__device__ __forceinline__ uint32_t col(uint32_t x, uint32_t lane4)
{
	return (((x >> 2) & (W - 1)) << 7) + lane4 + (x & 3u);
}

// n <= MAXB (16 or 32) bytes from registers into the lane's out ring at position d.  maxn = warp-uniform bound on n.
// A ring column belongs to one lane, so partial words need no byte stores: the first word is merged with the
// bytes already in front of d (read-modify-write), every further word is stored whole.  Bytes behind d + n in the
// last word are clobbered unless KEEP_TAIL: behind a literal run nothing has been written yet (a match piece
// whose literals were written first -- the deferred copy -- must keep them).
template <bool KEEP_TAIL, uint32_t MAXB>
__device__ __forceinline__ void store_bytes(uint8_t *outb, uint32_t lane4, const Bytes36 &D, uint32_t d, uint32_t n, uint32_t maxn,
					    bool active)
{
	if (!active) return;
	constexpr uint32_t M = OWW * 128 - 1;
	const uint32_t s8 = (d & 3u) * 8u;                    // bit position of byte d in its word
	const uint32_t nw = ((d & 3u) + n + 3u) >> 2;         // words touched, 1..MAXB / 4 + 1
	const uint32_t u0 = ((d >> 2) << 7) | lane4;
	uint32_t *first = reinterpret_cast<uint32_t *>(outb + (u0 & M));
	uint32_t *lastp = reinterpret_cast<uint32_t *>(outb + ((u0 + (nw - 1) * 128) & M));
	const uint32_t end8 = ((d + n) & 3u) * 8u;            // valid bits of the last word (0 = all 32)
	uint32_t old_last = 0;
	if (KEEP_TAIL && end8) old_last = *lastp;
	// word 0: payload shifted up to byte d, the bytes in front of d kept
	const uint32_t old0 = *first;
	const uint32_t low = (1u << s8) - 1u;               // s8 is 0, 8, 16 or 24
	*first = (D.w[0] << s8) | (old0 & low);
	// (the warp-uniform guard covers two words at a time: half the branches)
#pragma unroll
	for (int jj = 1; jj < static_cast<int>(MAXB / 4 + 1); jj += 2) {
		if (static_cast<uint32_t>(jj * 4) < maxn + 4u) {
#pragma unroll
			for (int j = jj; j < jj + 2; j++)
				if (static_cast<uint32_t>(j) < nw)
					*reinterpret_cast<uint32_t *>(outb + ((u0 + j * 128) & M)) = __funnelshift_l(D.w[j - 1], D.w[j], s8);
		}
	}
	if (KEEP_TAIL && end8) {
		const uint32_t keep = 0xffffffffu << end8;        // bytes behind d + n
		*lastp = (*lastp & ~keep) | (old_last & keep);
	}
}

// n <= MAXB (16 or 32) bytes starting at position s of a lane's ring column into registers (byte 0 = position s;
// words 0 .. MAXB / 4 of D are written).
template <uint32_t W, uint32_t MAXB>
__device__ __forceinline__ void fetch_col(Bytes36 &D, const uint8_t *ringb, uint32_t lane4, uint32_t s, uint32_t n, uint32_t maxn,
					  bool active)
{
	if (!active) return;
	constexpr uint32_t M = W * 128 - 1;
	const uint32_t bs = (s & 3u) * 8u;
	const uint32_t u0 = ((s >> 2) << 7) | lane4;
	const uint32_t nw = (n + (s & 3u) + 3u) >> 2;   // words that hold payload
	constexpr int NX = MAXB / 4 + 2;   // words that can hold payload, + one for the shift
	uint32_t x[NX];
#pragma unroll
	for (int j = 0; j < NX; j++) x[j] = 0;
#pragma unroll
	for (int jj = 0; jj < NX; jj += 2) {
		if (static_cast<uint32_t>(jj * 4) < maxn + 4u) {
#pragma unroll
			for (int j = jj; j < jj + 2; j++)
				if (static_cast<uint32_t>(j) < nw) x[j] = *reinterpret_cast<const uint32_t *>(ringb + ((u0 + j * 128) & M));
		}
	}
#pragma unroll
	for (int j = 0; j < NX - 1; j++) D.w[j] = __funnelshift_r(x[j], x[j + 1], bs);
}

// 48 aligned bytes A|B|C whose payload starts at byte m -> registers (byte 0 = payload byte 0).
__device__ __forceinline__ void align48(Bytes36 &D, const uint4 &A, const uint4 &B, const uint4 &C, uint32_t m)
{
	uint32_t x0 = A.x, x1 = A.y, x2 = A.z, x3 = A.w, x4 = B.x, x5 = B.y, x6 = B.z, x7 = B.w, x8 = C.x, x9 = C.y, x10 = C.z,
		 x11 = C.w;
	if (m & 8) { x0 = x2; x1 = x3; x2 = x4; x3 = x5; x4 = x6; x5 = x7; x6 = x8; x7 = x9; x8 = x10; x9 = x11; x10 = 0; }
	if (m & 4) { x0 = x1; x1 = x2; x2 = x3; x3 = x4; x4 = x5; x5 = x6; x6 = x7; x7 = x8; x8 = x9; x9 = x10; }
	const uint32_t bs = (m & 3u) * 8u;
	D.w[0] = __funnelshift_r(x0, x1, bs); D.w[1] = __funnelshift_r(x1, x2, bs); D.w[2] = __funnelshift_r(x2, x3, bs);
	D.w[3] = __funnelshift_r(x3, x4, bs); D.w[4] = __funnelshift_r(x4, x5, bs); D.w[5] = __funnelshift_r(x5, x6, bs);
	D.w[6] = __funnelshift_r(x6, x7, bs); D.w[7] = __funnelshift_r(x7, x8, bs); D.w[8] = __funnelshift_r(x8, x9, bs);
}

// Old match source, two halves one trip apart: issue the aligned 16-byte granules around [p, p + n) into the
// lane's staging slot with cp.async (no registers held while the L2 / DRAM round trip is in flight) ...
__device__ __forceinline__ void stage_issue(uint4 *slot, const uint8_t *p, uint32_t n)
{
	const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p) & 15u);
	const uint8_t *base = p - m;
	cp_async16(slot, base);
	if (m + n > 16) cp_async16(slot + 1, base + 16);
	if (m + n > 32) cp_async16(slot + 2, base + 32);
}
// ... and pick them up after cp.async.wait_group in the next trip.
__device__ __forceinline__ void stage_take(Bytes36 &D, const uint4 *slot, const uint8_t *p, uint32_t n, bool active)
{
	if (!active) return;
	const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p) & 15u);
	const uint4 A = slot[0];
	uint4 B = make_uint4(0, 0, 0, 0), C = make_uint4(0, 0, 0, 0);
	if (m + n > 16) B = slot[1];
	if (m + n > 32) C = slot[2];
	align48(D, A, B, C, m);
}

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
	const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem));
}

// Four bytes at position x of the lane's in ring (byte 0 = position x): two word reads and a funnel shift.
__device__ __forceinline__ uint32_t peek32(const uint8_t *inb, uint32_t lane4, uint32_t x)
{
	constexpr uint32_t M = IWW * 128 - 1;
	const uint32_t u = (((x >> 2) << 7) | lane4) & M;
	const uint32_t w0 = *reinterpret_cast<const uint32_t *>(inb + u);
	const uint32_t w1 = *reinterpret_cast<const uint32_t *>(inb + ((u + 128) & M));
	return __funnelshift_r(w0, w1, (x & 3u) * 8u);
}

// One 16-byte stripe at position x of the lane's in ring into the four XXH32 accumulators (Rot_Mul, :982-985).
__device__ __forceinline__ void hash_stripe(const uint8_t *inb, uint32_t lane4, uint32_t x, uint32_t &a0, uint32_t &a1, uint32_t &a2,
					    uint32_t &a3)
{
	constexpr uint32_t M = IWW * 128 - 1;
	const uint32_t u = ((x >> 2) << 7) | lane4;
	const uint32_t bs = (x & 3u) * 8u;
	const uint32_t w0 = *reinterpret_cast<const uint32_t *>(inb + (u & M));
	const uint32_t w1 = *reinterpret_cast<const uint32_t *>(inb + ((u + 128) & M));
	const uint32_t w2 = *reinterpret_cast<const uint32_t *>(inb + ((u + 256) & M));
	const uint32_t w3 = *reinterpret_cast<const uint32_t *>(inb + ((u + 384) & M));
	const uint32_t w4 = *reinterpret_cast<const uint32_t *>(inb + ((u + 512) & M));
	a0 = xxh_round(a0, __funnelshift_r(w0, w1, bs));
	a1 = xxh_round(a1, __funnelshift_r(w1, w2, bs));
	a2 = xxh_round(a2, __funnelshift_r(w2, w3, bs));
	a3 = xxh_round(a3, __funnelshift_r(w3, w4, bs));
}

// Process_Variable_Length (lib/lz4ada.adb:724-735) for a 15 nibble: e1 / e2 = the first two extension bytes
// (already read from the in ring), anything longer (>= 525) continues in global memory.  q = position of the
// first extension byte on entry, of the byte after the last one on return.  false = the block ends inside.
__device__ __forceinline__ bool length_ext(uint32_t e1, uint32_t e2, const uint8_t *sbase, uint32_t &q, uint32_t a_end, uint32_t &len)
{
	if (q >= a_end) return false;
	q++;
	len += e1;
	if (e1 != 255) return true;
	if (q >= a_end) return false;
	q++;
	len += e2;
	if (e2 != 255) return true;
	uint32_t b;
	do {
		if (q >= a_end) return false;
		b = __ldg(sbase + q);
		q++;
		len += b;
	} while (b == 255 && len < 0x40000000u);
	return true;
}

// One warp: lanes take blocks from *counter until it reaches n_blocks.
__device__ __forceinline__ void decode_lanes(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
					     const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status,
					     uint32_t *counter, WarpMem &wm, int lane)
{
	uint8_t *inb = reinterpret_cast<uint8_t *>(&wm.in_w[0][0]);
	uint8_t *outb = reinterpret_cast<uint8_t *>(&wm.out_w[0][0]);
	const uint32_t lane4 = static_cast<uint32_t>(lane) * 4u;

	// ---- lane state.  Input positions are in "aligned stream coordinates": position of a payload byte
	// = its offset in the block + mis, so that position 16 c is the start of the c-th aligned 16-byte
	// chunk in global memory.  Output positions likewise: offset in the block's output + ophase.
	uint32_t state = L_IDLE, sq = S_TOKEN, blk = 0;
	const uint8_t *sbase = src;    // 16-byte aligned: payload byte at position x is sbase[x]
	uint8_t *obase = dst;          // 16-byte aligned: output byte at position x is obase[x]
	uint32_t a_cur = 0, a_end = 0, a_loaded = 0;
	uint32_t p_cur = 0, p_start = 0, p_cap = 0, p_flushed = 0, ring_lo = 0;
	uint32_t declared = 0;
	uint32_t rem_lit = 0, rem_ml = 0, dist = 0, mln = 0;   // the sequence in progress; a parked giant is the literal
	                                                       // run rem_lit (sq == S_LIT) or the match rem_ml / dist
	// fused block checksum (XXHash32.Process, lib/lz4ada.adb:979-991): four accumulators, next stripe at h_pos
	uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, h_pos = 0, a_beg = 0, h_mode = H_NONE;
	uint32_t rf = 0;                                       // bytes of the in ring refill in flight (0 or 16)
	// the match piece in flight: set up (and, for an old source, requested) at the end of one trip, copied at the
	// end of the next one, so that the memory round trip overlaps that trip's parsing, literals and flush.
	// p_cur already counts it; its n bytes at pend_dst are a hole until then.
	uint32_t pend_n = 0, pend_dst = 0, pend_src = 0;
	bool pend_far = false;
	bool exhausted = false;
	uint32_t idle_trips = 0;   // trips in which no lane made progress: a safety net, never reached on valid state

	for (;;) {
		// ================= fetch new blocks when enough lanes are idle =================
		const uint32_t idle = __ballot_sync(FULL_MASK, state == L_IDLE);
		if (idle && !exhausted && (__popc(idle) >= 8 || !__any_sync(FULL_MASK, state == L_RUN))) {
			// (eight at a time: a refill stalls the warp for one round of block hashing)
			uint32_t base = 0;
			const uint32_t want = __popc(idle);
			if (lane == 0) base = atomicAdd(counter, want);
			base = __shfl_sync(FULL_MASK, base, 0);
			if (base + want >= n_blocks) exhausted = true;
			uint32_t hflags = 0, hn = 0;
			const uint8_t *hs = src;
			bool fresh = false;
			if (state == L_IDLE) {
				const uint32_t b = base + __popc(idle & ((1u << lane) - 1u));
				if (b < n_blocks) {
					const lz4b200_blk_desc d = desc[b];
					if (!(d.flags & LZ4B200_BLK_NOT_K1)) {
						blk = b;
						fresh = true;
						hflags = d.flags;
						hn = d.src_len;
						hs = src + d.src_off;
						const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(hs) & 15u);
						uint8_t *og = dst + d.dst_off;
						const uint32_t oph = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(og) & 15u);
						sbase = hs - mis;
						obase = og - oph;
						a_cur = mis;
						a_end = mis + d.src_len;
						a_loaded = 0;
						p_start = p_cur = oph;
						p_cap = oph + d.dst_cap;
						p_flushed = 0;   // chunk 0 is partial when oph != 0: flushed bytewise
						ring_lo = oph;
						declared = 0;
						a_beg = h_pos = mis;
						acc0 = PRIME_1 + PRIME_2; acc1 = PRIME_2; acc2 = 0; acc3 = 0u - PRIME_1;   // Reset, :932-940
						sq = S_TOKEN;
						rem_lit = rem_ml = 0;
						pend_n = 0;
						// positions are 32-bit with headroom; odd blocks go to the exact routine
						const bool plain = !(d.flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_HASH_ONLY)) && d.dst_cap < 0x7fff0000u &&
								   d.src_len < 0x7fff0000u;
						state = plain ? L_RUN : L_EXACT;
					}
				}
			}
			if (fresh) h_mode = H_NONE;
			if (fresh && (hflags & LZ4B200_BLK_HAS_CHECKSUM) && state == L_RUN) {
				const uint8_t *t = hs + hn;
				declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
				h_mode = H_INLINE;
			}
			__syncwarp();
		}
		if (__all_sync(FULL_MASK, state == L_IDLE)) {
			if (exhausted) break;
			continue;
		}

		// ================= parked work that needs the whole warp =================
		uint32_t coop = __ballot_sync(FULL_MASK, state == L_GIANT || state == L_EXACT || state == L_HASH);
		if (coop) {
			// (rare) the refills in flight must have landed before a lane restarts its in ring
			cp_async_wait<0>();
			a_loaded += rf;
			rf = 0;
		}
		while (coop) {
			const int j = __ffs(coop) - 1;
			coop &= coop - 1;
			const uint32_t stj = __shfl_sync(FULL_MASK, state, j);
			const uint32_t bj = __shfl_sync(FULL_MASK, blk, j);
			if (stj == L_EXACT) {
				const lz4b200_blk_desc d = desc[bj];
				process_block<false>(src, dst + d.dst_off, d, d.dst_cap, d.hist_avail, status + bj, lane);
				if (lane == j) state = L_IDLE;
				__syncwarp();
				continue;
			}
			const uint8_t *sb = shfl_cptr(sbase, j);
			if (stj == L_HASH) {
				// the block's checksum over again (a giant made the lane skip parts of its in ring): quad 0 hashes,
				// the lane compares and reports
				const uint32_t begj = __shfl_sync(FULL_MASK, a_beg, j), endj = __shfl_sync(FULL_MASK, a_end, j);
				const uint32_t h = __shfl_sync(FULL_MASK, quad_xxh32_prologue(sb + begj, lane < 4 ? endj - begj : 0u, lane), 0);
				if (lane == j) {
					lz4b200_blk_status *st = status + blk;
					const bool okay = h == declared;
					st->code = okay ? LZ4B200_ST_OK : LZ4B200_ST_BLOCK_CHECKSUM;
					st->out_len = okay ? p_cur - p_start : 0u;
					st->err_pos = 0;
					st->aux = 0;
					st->xxh32_computed = h;
					st->xxh32_declared = declared;
					state = L_IDLE;
				}
				__syncwarp();
				continue;
			}
			// one giant copy of lane j, straight in global memory (the lane has flushed its ring)
			uint8_t *ob = shfl_ptr(obase, j);
			const uint32_t g_len = sq == S_LIT ? rem_lit : rem_ml, g_off = sq == S_LIT ? 0u : dist;
			const uint32_t lenj = __shfl_sync(FULL_MASK, g_len, j), offj = __shfl_sync(FULL_MASK, g_off, j);
			const uint32_t pj = __shfl_sync(FULL_MASK, p_cur, j), aj = __shfl_sync(FULL_MASK, a_cur, j);
			__syncwarp();
			if (offj == 0) warp_copy<true>(ob + pj, sb + aj, lenj, lane);
			else match_copy(ob + pj, offj, lenj, lane);
			__syncwarp();
			if (lane == j) {
				p_cur += g_len;
				if (g_off == 0) {
					a_cur += g_len;
					rem_lit = 0;
					sq = S_OFF;
					a_loaded = a_cur & ~15u;   // the in ring starts over at the chunk under a_cur
				} else {
					rem_ml = 0;
					sq = S_TOKEN;
				}
				// re-seed: the out ring holds nothing but the (partial) chunk under p_cur
				p_flushed = p_cur & ~15u;
				ring_lo = p_flushed > p_start ? p_flushed : p_start;
				if (p_cur > p_flushed) {
					const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(obase + p_flushed));
					const uint32_t u0 = ((p_flushed >> 2) << 7) | lane4;
					constexpr uint32_t M = OWW * 128 - 1;
					*reinterpret_cast<uint32_t *>(outb + ((u0 + 0) & M)) = v.x;
					*reinterpret_cast<uint32_t *>(outb + ((u0 + 128) & M)) = v.y;
					*reinterpret_cast<uint32_t *>(outb + ((u0 + 256) & M)) = v.z;
					*reinterpret_cast<uint32_t *>(outb + ((u0 + 384) & M)) = v.w;
				}
				state = L_RUN;
			}
			__syncwarp();
		}

		const bool run = state == L_RUN;
		bool bad = false, giant = false, progressed = false;
		uint4 *slot = &wm.stage[lane][0];
		const uint32_t have = a_loaded < a_end ? a_loaded : a_end;   // payload bytes are valid below this
		const bool all_in = have == a_end;

		Bytes36 D;   // written by whichever fetch precedes a store; a lane that fetches nothing stores nothing
		// ================= stage 1: token (Decompress_Sequence, lib/lz4ada.adb:737-750) =================
		if (run && sq == S_TOKEN) {
			const uint32_t avail = have > a_cur ? have - a_cur : 0u;
			if (a_cur >= a_end) {
				state = L_FINISH;   // the block ends after a match, or is empty
			} else if (avail >= 3 || all_in) {
				const uint32_t pk = peek32(inb, lane4, a_cur);   // token and two possible extension bytes
				const uint32_t tk = pk & 255u;
				uint32_t lit = tk >> 4;
				mln = tk & 15u;
				uint32_t q = a_cur + 1;
				if (lit == 15 && !length_ext((pk >> 8) & 255u, (pk >> 16) & 255u, sbase, q, a_end, lit)) bad = true;
				if (!bad && (lit > a_end - q || lit > p_cap - p_cur)) bad = true;
				if (!bad) {
					a_cur = q;
					rem_lit = lit;
					sq = lit ? S_LIT : S_OFF;
					progressed = true;
					if (lit >= GIANT) giant = true;
				}
			}
		}
		// ================= stage 2: up to 16 literal bytes, in ring -> registers -> out ring (:790-824) =================
		{
			const uint32_t avail = have > a_cur ? have - a_cur : 0u;
			uint32_t n = 0;
			if (state == L_RUN && !bad && !giant && sq == S_LIT) {
				n = rem_lit < LIT_PIECE ? rem_lit : LIT_PIECE;
				n = n < avail ? n : avail;
			}
			const uint32_t maxn = __reduce_max_sync(FULL_MASK, n);
			if (maxn) {
				fetch_col<IWW, LIT_PIECE>(D, inb, lane4, a_cur, n, maxn, n != 0);
				store_bytes<false, LIT_PIECE>(outb, lane4, D, p_cur, n, maxn, n != 0);
				if (n) {
					a_cur += n;
					p_cur += n;
					rem_lit -= n;
					if (rem_lit == 0) sq = S_OFF;
					progressed = true;
				}
			}
		}
		// ================= stage 3: offset and match length (:752-777) =================
		if (state == L_RUN && !bad && !giant && sq == S_OFF) {
			const uint32_t avail = have > a_cur ? have - a_cur : 0u;
			if (a_cur >= a_end) {
				// final literal-only sequence; a match nibble here is an error (:752-764)
				if (mln) bad = true;
				else state = L_FINISH;
			} else if (avail >= 4 || all_in) {
				if (a_cur + 2 > a_end) {
					bad = true;
				} else {
					const uint32_t pk = peek32(inb, lane4, a_cur);   // offset and two possible extension bytes
					const uint32_t off = pk & 0xffffu;
					uint32_t q = a_cur + 2, ml = mln;
					if (ml == 15 && !length_ext((pk >> 16) & 255u, pk >> 24, sbase, q, a_end, ml)) bad = true;
					ml += 4;
					if (!bad && (off == 0 || off > p_cur - p_start || ml > p_cap - p_cur)) bad = true;
					if (!bad) {
						a_cur = q;
						rem_ml = ml;
						dist = off;
						sq = S_MATCH;
						progressed = true;
						if (ml >= GIANT) giant = true;
					}
				}
			}
		}
		if (bad) state = L_EXACT;   // the exact routine re-decodes the block and reports the error

		// ================= flush the out ring: 32 bytes (two aligned 16-byte stores) at a time =================
		auto flush16 = [&]() {
			const uint32_t u0 = ((p_flushed >> 2) << 7) | lane4;
			constexpr uint32_t M = OWW * 128 - 1;
			uint4 v;
			v.x = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 0) & M));
			v.y = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 128) & M));
			v.z = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 256) & M));
			v.w = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 384) & M));
			if (p_flushed >= p_start) {
				*reinterpret_cast<uint4 *>(obase + p_flushed) = v;
			} else {
				// the block's first chunk starts in the middle of a 16-byte granule: bytes only
				const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
				for (int k = 0; k < 16; k++)
					if (p_flushed + k >= p_start) obase[p_flushed + k] = static_cast<uint8_t>(ws[k >> 2] >> (8 * (k & 3)));
			}
			p_flushed += 16;
		};
		const uint32_t p_solid = pend_n ? pend_dst : p_cur;   // output is complete below this position
#pragma unroll 1
		for (int rep = 0; rep < 3; rep++) {
			const bool fl = (state == L_RUN || state == L_FINISH) && p_solid - p_flushed >= 32u && p_solid > p_flushed;
			if (!__any_sync(FULL_MASK, fl)) break;
			if (fl) {
				if (p_flushed < p_start) {
					flush16();
				} else {
					const uint32_t u0 = ((p_flushed >> 2) << 7) | lane4;
					constexpr uint32_t M = OWW * 128 - 1;
					uint4 v, w;
					v.x = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 0) & M));
					v.y = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 128) & M));
					v.z = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 256) & M));
					v.w = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 384) & M));
					w.x = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 512) & M));
					w.y = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 640) & M));
					w.z = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 768) & M));
					w.w = *reinterpret_cast<const uint32_t *>(outb + ((u0 + 896) & M));
					*reinterpret_cast<uint4 *>(obase + p_flushed) = v;
					*reinterpret_cast<uint4 *>(obase + p_flushed + 16) = w;
					p_flushed += 32;
				}
			}
		}
		// ================= everything requested in the last trip has landed =================
		cp_async_wait<0>();
		a_loaded += rf;
		rf = 0;
		// ================= fused block checksum: one stripe per trip, straight out of the in ring =================
		{
			const bool hs_ready = state == L_RUN && h_mode == H_INLINE && h_pos + 16u <= a_end && a_loaded >= h_pos + 16u;
			if (__any_sync(FULL_MASK, hs_ready)) {
				if (hs_ready) {
					hash_stripe(inb, lane4, h_pos, acc0, acc1, acc2, acc3);
					h_pos += 16;
				}
			}
		}
		// ================= stage 4: the match piece in flight, <= 32 bytes (Output_With_History, :845-904) =================
		{
			const uint32_t n = (state == L_EXACT) ? 0u : pend_n;
			const uint32_t maxn = __reduce_max_sync(FULL_MASK, n);
			if (maxn) {
				const bool is_far = n != 0 && pend_far, is_near = n != 0 && !pend_far;
				if (__any_sync(FULL_MASK, is_far)) stage_take(D, slot, obase + pend_src, n, is_far);
				// (the words to read follow from the longest *young* piece, usually much shorter than the longest piece)
				const uint32_t maxnear = __reduce_max_sync(FULL_MASK, is_near ? n : 0u);
				if (maxnear) fetch_col<OWW, ML_PIECE>(D, outb, lane4, pend_src, n, maxnear, is_near);
				store_bytes<true, ML_PIECE>(outb, lane4, D, pend_dst, n, maxn, n != 0);
				if (n) progressed = true;
			}
			pend_n = 0;
		}
		// ---- a lane that parks or finishes writes what is left, whole chunks and the partial one:
		//      global memory is then complete ----
		if ((state == L_RUN && giant) || (state == L_FINISH && p_cur - p_flushed < 32u)) {
			while ((p_cur & ~15u) > p_flushed) flush16();
			const uint32_t base16 = p_flushed;
			if (p_cur > base16) {
				const uint32_t u0 = ((base16 >> 2) << 7) | lane4;
				constexpr uint32_t M = OWW * 128 - 1;
				uint32_t ws[4];
#pragma unroll
				for (int k = 0; k < 4; k++) ws[k] = *reinterpret_cast<const uint32_t *>(outb + ((u0 + k * 128) & M));
#pragma unroll
				for (int k = 0; k < 16; k++)
					if (base16 + k >= p_start && base16 + k < p_cur) obase[base16 + k] = static_cast<uint8_t>(ws[k >> 2] >> (8 * (k & 3)));
			}
			if (state == L_FINISH) {
				uint32_t computed = 0;
				bool okay = true;
				if (h_mode == H_INLINE) {
					// the stripes the hash is still behind (their chunks are kept in the in ring), then Final (:993-1017)
					while (h_pos + 16u <= a_end) {
						hash_stripe(inb, lane4, h_pos, acc0, acc1, acc2, acc3);
						h_pos += 16;
					}
					computed = xxh_finish<true>(acc0, acc1, acc2, acc3, a_end - a_beg, sbase + h_pos, a_end - h_pos);
					okay = computed == declared;
				}
				if (h_mode == H_REDO) {
					state = L_HASH;   // needs a quad of the warp: next trip
				} else {
					lz4b200_blk_status *st = status + blk;
					st->code = okay ? LZ4B200_ST_OK : LZ4B200_ST_BLOCK_CHECKSUM;   // :672-676, :702
					st->out_len = okay ? p_cur - p_start : 0u;
					st->err_pos = 0;
					st->aux = 0;
					st->xxh32_computed = computed;
					st->xxh32_declared = declared;
					state = L_IDLE;
				}
			} else {
				if (h_mode == H_INLINE) h_mode = H_REDO;   // the giant's bytes never pass through the in ring
				state = L_GIANT;
			}
			progressed = true;
		}
		// ================= requests for the next trip (cp.async: no registers wait for them) =================
		{
			// set up the next match piece (copied at the end of the next trip); an old source is requested now
			if (state == L_RUN && sq == S_MATCH && rem_ml != 0) {
				uint32_t n = rem_ml < ML_PIECE ? rem_ml : ML_PIECE;
				n = n < dist ? n : dist;   // a piece never overlaps its own source
				const uint32_t src_s = p_cur - dist;
				// what the ring still holds when the piece is copied: the next trip may write up to 16 literal bytes
				// behind it first
				const uint32_t hi = p_cur + n + LIT_PIECE + 4u;   // + the bytes a literal store clobbers behind its end
				const uint32_t lo_wr = hi > OUT_BYTES ? hi - OUT_BYTES : 0u;
				const uint32_t near_lo = ring_lo > lo_wr ? ring_lo : lo_wr;
				pend_far = src_s < near_lo;
				if (pend_far) {
					// read from global memory, up to the flush frontier (the rest follows as a young source)
					if (p_flushed - src_s < n) n = p_flushed - src_s;
					stage_issue(slot, obase + src_s, n);
				}
				pend_n = n;
				pend_dst = p_cur;
				pend_src = src_s;
				p_cur += n;
				rem_ml -= n;
				if (dist < ML_PIECE && n == dist) dist <<= 1;   // the pattern has doubled
				if (rem_ml == 0) sq = S_TOKEN;
				progressed = true;
			}
			// in ring: one aligned chunk per trip when there is room
			// (a length read from global memory may have carried a_cur past the loaded chunks: skip them)
			if (a_loaded < (a_cur & ~15u)) a_loaded = a_cur & ~15u;
			const uint32_t a_end16 = (a_end + 15u) & ~15u;
			// the oldest byte still needed: the decoder's position, or the checksum's if it is behind
			const uint32_t a_tail = (h_mode == H_INLINE && h_pos < a_cur ? h_pos : a_cur) & ~15u;
			const bool room = state == L_RUN && a_loaded < a_end16 && a_loaded + 16u - a_tail <= IN_BYTES;
			if (room) {
				const uint8_t *g = sbase + a_loaded;
				const uint32_t u0 = ((a_loaded >> 2) << 7) | lane4;
				constexpr uint32_t M = IWW * 128 - 1;
				cp_async4(inb + ((u0 + 0) & M), g);
				cp_async4(inb + ((u0 + 128) & M), g + 4);
				cp_async4(inb + ((u0 + 256) & M), g + 8);
				cp_async4(inb + ((u0 + 384) & M), g + 12);
				rf = 16;
				progressed = true;
			}
			cp_async_commit();
		}
		// safety net: a state the lock-step machine cannot leave must not hang the device -- hand the blocks to the
		// exact routine (a lane makes progress whenever it consumes input, writes output, parks or finishes)
		{
			const bool moving = progressed || state != L_RUN;
			idle_trips = __any_sync(FULL_MASK, moving) ? 0u : idle_trips + 1u;
			if (idle_trips > 4096u) {
				if (state == L_RUN) state = L_EXACT;
				idle_trips = 0;
			}
		}
		__syncwarp();
	}
}

}  // namespace v5
}  // namespace lz4b200
