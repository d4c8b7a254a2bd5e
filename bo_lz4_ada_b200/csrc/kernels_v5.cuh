// kernels_v5.cuh -- K1 fifth generation: one LANE per block, 32 blocks in lock-step per warp.
//
// v2 / v4 give a block to a warp and a sequence to a lane.  ncu (profiles/r01_v4_*): 37 warp
// instructions per LZ4 sequence, 15-20 of 32 lanes active, because a sequence is ~10 bytes of work and
// everything around it (scans, dependency search, the serial hand-over of dependent matches) is
// per-batch overhead.  v5 turns the decomposition around, the way a CPU decodes: every lane walks its
// own block, one sequence per trip of a warp-uniform loop, so all 32 lanes do useful work in every
// instruction and nothing is ever exchanged between lanes:
//
//   in ring   128 bytes of the lane's compressed stream, refilled with one aligned 16-byte global load
//             per trip when there is room
//   out ring  the lane's last 256 output bytes.  Literals and matches are assembled here; complete
//             16-byte chunks go to global memory with one aligned 16-byte store each (full sectors)
//   layout    word w of lane L lives at shared-memory word w * 32 + L: lane L only ever touches bank L,
//             so 32 lanes reading 32 unrelated positions is one conflict-free wavefront
//   matches   source younger than ~200 bytes: out ring; older: three aligned 16-byte global loads
//             (flushed long ago); both land in registers, then one common routine stores <= 32 bytes
//   long ops  literal runs > 16, matches > 32, self-overlapping matches, length extensions of 255:
//             the lane parks and the whole warp does that one sequence in global memory with the v1
//             warp copies (RLE blocks are one such sequence), then the lane re-seeds its rings
//
// Lanes fetch new blocks from a global counter when enough of them are idle, so unequal blocks do not
// leave lanes parked.  Block checksums (Check_Checksum, lib/lz4ada.adb:698-707): the quads of the warp
// hash the newly fetched blocks, eight at a time, before they are decoded.
// Anything unusual goes to process_block (exact error semantics of lib/lz4ada.adb:716-904).
#pragma once

#include "kernels_v2.cuh"

namespace lz4b200 {
namespace v5 {

constexpr int WARPS = 4;                  // per CTA
constexpr uint32_t IWW = 32;              // in ring: words per lane (128 bytes)
constexpr uint32_t OWW = 64;              // out ring: words per lane (256 bytes)
constexpr uint32_t IN_BYTES = IWW * 4, OUT_BYTES = OWW * 4;
constexpr uint32_t LIT_SHORT = 16;        // literal runs up to this length stay in the lane
constexpr uint32_t ML_SHORT = 32;         // matches up to this length stay in the lane
constexpr uint32_t NEED = 24;             // compressed bytes a short sequence may touch: 1 + 1 + 16 + 2 + 1, rounded up

struct __align__(16) WarpMem {
	uint32_t in_w[IWW][32];
	uint32_t out_w[OWW][32];
};

enum : uint32_t { L_IDLE = 0, L_RUN = 1, L_LONG = 2, L_EXACT = 3, L_FINISH = 4 };

// Byte address of position x (aligned stream coordinates) in a lane's column of a ring of W words.
template <uint32_t W>
__device__ __forceinline__ uint32_t col(uint32_t x, uint32_t lane4)
{
	return (((x >> 2) & (W - 1)) << 7) + lane4 + (x & 3u);
}

struct Bytes36 { uint32_t w[9]; };   // up to 32 payload bytes starting at byte 0 (+ one word of slack for shifting)

// n <= 32 bytes from registers into the lane's out ring at position d.  maxn = warp-uniform bound on n.
__device__ __forceinline__ void store_bytes(uint8_t *outb, uint32_t lane4, const Bytes36 &D, uint32_t d, uint32_t n, uint32_t maxn,
					    bool active)
{
	if (!active) return;
	const uint32_t hb0 = (4u - (d & 3u)) & 3u;
	const uint32_t hb = hb0 < n ? hb0 : n;   // head bytes up to the next word boundary
	uint8_t *hp = outb + col<OWW>(d, lane4);
	if (hb > 0) hp[0] = static_cast<uint8_t>(D.w[0]);
	if (hb > 1) hp[1] = static_cast<uint8_t>(D.w[0] >> 8);
	if (hb > 2) hp[2] = static_cast<uint8_t>(D.w[0] >> 16);
	const uint32_t hs = hb * 8;
	const uint32_t nwords = (n - hb) >> 2;
	const uint32_t w0 = (d + hb) >> 2;   // first full word
	uint32_t tail = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		if (static_cast<uint32_t>(j * 4) < maxn) {
			const uint32_t v = __funnelshift_r(D.w[j], D.w[j + 1], hs);
			if (static_cast<uint32_t>(j) < nwords)
				*reinterpret_cast<uint32_t *>(outb + (((w0 + j) & (OWW - 1)) << 7) + lane4) = v;
			if (static_cast<uint32_t>(j) == nwords) tail = v;
		}
	}
	const uint32_t tb = hb + 4u * nwords;
	if (tb < n) {
		uint8_t *tp = outb + (((w0 + nwords) & (OWW - 1)) << 7) + lane4;
		tp[0] = static_cast<uint8_t>(tail);
		if (tb + 1 < n) tp[1] = static_cast<uint8_t>(tail >> 8);
		if (tb + 2 < n) tp[2] = static_cast<uint8_t>(tail >> 16);
	}
}

// n <= 32 bytes starting at position s of a lane's ring column into registers (byte 0 = position s).
template <uint32_t W>
__device__ __forceinline__ void fetch_col(Bytes36 &D, const uint8_t *ringb, uint32_t lane4, uint32_t s, uint32_t n, uint32_t maxn,
					  bool active)
{
	if (!active) return;
	const uint32_t ws = s >> 2, bs = (s & 3u) * 8u;
	uint32_t x[10];
#pragma unroll
	for (int j = 0; j < 10; j++) {
		x[j] = 0;
		if (static_cast<uint32_t>(j * 4) < maxn + 7u && static_cast<uint32_t>(j * 4) < n + (s & 3u))
			x[j] = *reinterpret_cast<const uint32_t *>(ringb + (((ws + j) & (W - 1)) << 7) + lane4);
	}
#pragma unroll
	for (int j = 0; j < 9; j++) D.w[j] = __funnelshift_r(x[j], x[j + 1], bs);
}

// n <= 32 bytes starting at global address p into registers (three aligned 16-byte loads).
__device__ __forceinline__ void fetch_global(Bytes36 &D, const uint8_t *p, uint32_t n, bool active)
{
	if (!active) return;
	const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p) & 15u);
	const uint4 *base = reinterpret_cast<const uint4 *>(p - m);
	const uint4 A = __ldcg(base);
	uint4 B = make_uint4(0, 0, 0, 0), C = make_uint4(0, 0, 0, 0);
	if (m + n > 16) B = __ldcg(base + 1);
	if (m + n > 32) C = __ldcg(base + 2);
	uint32_t x0 = A.x, x1 = A.y, x2 = A.z, x3 = A.w, x4 = B.x, x5 = B.y, x6 = B.z, x7 = B.w, x8 = C.x, x9 = C.y, x10 = C.z,
		 x11 = C.w;
	if (m & 8) { x0 = x2; x1 = x3; x2 = x4; x3 = x5; x4 = x6; x5 = x7; x6 = x8; x7 = x9; x8 = x10; x9 = x11; x10 = 0; }
	if (m & 4) { x0 = x1; x1 = x2; x2 = x3; x3 = x4; x4 = x5; x5 = x6; x6 = x7; x7 = x8; x8 = x9; x9 = x10; }
	const uint32_t bs = (m & 3u) * 8u;
	D.w[0] = __funnelshift_r(x0, x1, bs); D.w[1] = __funnelshift_r(x1, x2, bs); D.w[2] = __funnelshift_r(x2, x3, bs);
	D.w[3] = __funnelshift_r(x3, x4, bs); D.w[4] = __funnelshift_r(x4, x5, bs); D.w[5] = __funnelshift_r(x5, x6, bs);
	D.w[6] = __funnelshift_r(x6, x7, bs); D.w[7] = __funnelshift_r(x7, x8, bs); D.w[8] = __funnelshift_r(x8, x9, bs);
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
	uint32_t state = L_IDLE, blk = 0;
	const uint8_t *sbase = src;    // 16-byte aligned: payload byte at position x is sbase[x]
	uint8_t *obase = dst;          // 16-byte aligned: output byte at position x is obase[x]
	uint32_t a_cur = 0, a_end = 0, a_loaded = 0, a_start = 0;
	uint32_t p_cur = 0, p_start = 0, p_cap = 0, p_flushed = 0, ring_lo = 0;
	uint32_t computed = 0, declared = 0;
	// a parked sequence (L_LONG)
	uint32_t lg_lit = 0, lg_q = 0, lg_ml = 0, lg_off = 0, lg_nx = 0;
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
					if (!(d.flags & LZ4B200_BLK_CHAINED)) {
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
						a_start = a_cur = mis;
						a_end = mis + d.src_len;
						a_loaded = 0;
						p_start = p_cur = oph;
						p_cap = oph + d.dst_cap;
						p_flushed = 0;   // chunk 0 is partial when oph != 0: flushed bytewise
						ring_lo = oph;
						computed = declared = 0;
						// positions are 32-bit with headroom; odd blocks go to the exact routine
						const bool plain = !(d.flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_HASH_ONLY)) && d.dst_cap < 0x7fff0000u &&
								   d.src_len < 0x7fff0000u;
						state = plain ? L_RUN : L_EXACT;
					}
				}
			}
			// ---- block checksums of the fresh blocks: quad q hashes the q-th of them, eight per round ----
			uint32_t todo = __ballot_sync(FULL_MASK, fresh && (hflags & LZ4B200_BLK_HAS_CHECKSUM) && state == L_RUN);
			while (todo) {
				// the q-th set bit of todo
				uint32_t mine = __fns(todo, 0, (lane >> 2) + 1);
				if (mine > 31) mine = 32;
				const int srcl = mine < 32 ? static_cast<int>(mine) : 0;
				const uint8_t *sq = shfl_cptr(hs, srcl);
				const uint32_t nq = __shfl_sync(FULL_MASK, hn, srcl);
				const uint32_t h = quad_xxh32_prologue(sq, mine < 32 ? nq : 0u, lane);
				// hand the digest to the owning lane: lane `mine` takes it from the first lane of quad q
				uint32_t taken = 0;
#pragma unroll
				for (int qq = 0; qq < 8; qq++) {
					const uint32_t owner = __shfl_sync(FULL_MASK, mine, qq * 4);
					const uint32_t hv = __shfl_sync(FULL_MASK, h, qq * 4);
					if (owner == static_cast<uint32_t>(lane)) computed = hv;
					if (owner < 32) taken |= 1u << owner;
				}
				todo &= ~taken;
			}
			if (fresh && (hflags & LZ4B200_BLK_HAS_CHECKSUM) && state == L_RUN) {
				const uint8_t *t = hs + hn;
				declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
				if (computed != declared) {
					// :672-676 -- verified before any decoding
					lz4b200_blk_status *st = status + blk;
					st->code = LZ4B200_ST_BLOCK_CHECKSUM;
					st->out_len = 0; st->err_pos = 0; st->aux = 0;
					st->xxh32_computed = computed; st->xxh32_declared = declared;
					state = L_IDLE;
				}
			}
			__syncwarp();
		}
		if (__all_sync(FULL_MASK, state == L_IDLE)) {
			if (exhausted) break;
			continue;
		}

		// ================= parked work that needs the whole warp =================
		uint32_t coop = __ballot_sync(FULL_MASK, state == L_LONG || state == L_EXACT);
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
			// one long sequence of lane j, straight in global memory (the lane has flushed its ring)
			const uint8_t *sb = shfl_cptr(sbase, j);
			uint8_t *ob = shfl_ptr(obase, j);
			const uint32_t litj = __shfl_sync(FULL_MASK, lg_lit, j), qj = __shfl_sync(FULL_MASK, lg_q, j);
			const uint32_t mlj = __shfl_sync(FULL_MASK, lg_ml, j), offj = __shfl_sync(FULL_MASK, lg_off, j);
			const uint32_t pj = __shfl_sync(FULL_MASK, p_cur, j);
			__syncwarp();
			if (litj) warp_copy<true>(ob + pj, sb + qj, litj, lane);
			if (mlj) match_copy(ob + pj + litj, offj, mlj, lane);
			__syncwarp();
			if (lane == j) {
				p_cur += lg_lit + lg_ml;
				a_cur = lg_nx;
				// re-seed: the out ring holds nothing but the (partial) chunk under p_cur
				p_flushed = p_cur & ~15u;
				ring_lo = p_flushed > p_start ? p_flushed : p_start;
				if (p_cur > p_flushed) {
					const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(obase + p_flushed));
					const uint32_t w0 = p_flushed >> 2;
					*reinterpret_cast<uint32_t *>(outb + (((w0 + 0) & (OWW - 1)) << 7) + lane4) = v.x;
					*reinterpret_cast<uint32_t *>(outb + (((w0 + 1) & (OWW - 1)) << 7) + lane4) = v.y;
					*reinterpret_cast<uint32_t *>(outb + (((w0 + 2) & (OWW - 1)) << 7) + lane4) = v.z;
					*reinterpret_cast<uint32_t *>(outb + (((w0 + 3) & (OWW - 1)) << 7) + lane4) = v.w;
				}
				a_loaded = a_cur & ~15u;   // the in ring starts over at the chunk under a_cur
				state = a_cur >= a_end ? L_FINISH : L_RUN;
			}
			__syncwarp();
		}

		// ================= refill the in ring: one aligned chunk per trip when there is room =================
		const bool run = state == L_RUN;
		{
			const uint32_t a_end16 = (a_end + 15u) & ~15u;
			const bool room = run && a_loaded < a_end16 && a_loaded + 16u - (a_cur & ~15u) <= IN_BYTES;
			if (room) {
				const uint4 v = __ldg(reinterpret_cast<const uint4 *>(sbase + a_loaded));
				const uint32_t w0 = a_loaded >> 2;
				*reinterpret_cast<uint32_t *>(inb + (((w0 + 0) & (IWW - 1)) << 7) + lane4) = v.x;
				*reinterpret_cast<uint32_t *>(inb + (((w0 + 1) & (IWW - 1)) << 7) + lane4) = v.y;
				*reinterpret_cast<uint32_t *>(inb + (((w0 + 2) & (IWW - 1)) << 7) + lane4) = v.z;
				*reinterpret_cast<uint32_t *>(inb + (((w0 + 3) & (IWW - 1)) << 7) + lane4) = v.w;
				a_loaded += 16;
			}
		}

		// ================= one sequence per running lane =================
		// (Decompress_Sequence, lib/lz4ada.adb:737-777)
		const uint32_t have = a_loaded < a_end ? a_loaded : a_end;   // payload bytes are valid below this
		bool go = run && a_cur < a_end && have > a_cur && (have - a_cur >= NEED || have == a_end);
		uint32_t lit = 0, ml = 0, off = 0, q = 0, nx = 0;
		bool park = false, bad = false;
		if (go) {
			const uint32_t tk = inb[col<IWW>(a_cur, lane4)];
			lit = tk >> 4;
			ml = tk & 15u;
			q = a_cur + 1;
			if (lit == 15) {
				if (q >= a_end) {
					bad = true;
				} else {
					const uint32_t b = inb[col<IWW>(q, lane4)];
					q++;
					lit += b;
					if (b == 255) park = true;
				}
			}
			const uint32_t e = q + lit;
			nx = e;
			if (!bad && !park) {
				if (lit > LIT_SHORT) {
					park = true;
				} else if (e > a_end) {
					bad = true;
				} else if (e == a_end) {
					// final literal-only sequence (:752-764)
					if (ml) bad = true;
					ml = 0;
				} else if (e + 2 > a_end) {
					bad = true;
				} else {
					off = inb[col<IWW>(e, lane4)] | (static_cast<uint32_t>(inb[col<IWW>(e + 1, lane4)]) << 8);
					nx = e + 2;
					if (ml == 15) {
						if (nx >= a_end) {
							bad = true;
						} else {
							const uint32_t b = inb[col<IWW>(nx, lane4)];
							nx++;
							ml += b;
							if (b == 255) park = true;
						}
					}
					ml += 4;
					if (ml > ML_SHORT || off < ml) park = true;
					if (off == 0 || off > p_cur + lit - p_start) bad = true;
				}
			}
			if (!bad && !park && lit + ml > p_cap - p_cur) bad = true;
			if (bad) park = false;
		}
		// a lane that must park parses the sequence in full first (extensions of any length, from global memory)
		if (go && park) {
			const uint8_t *s = sbase;
			uint32_t x = a_cur;
			const uint32_t tk = __ldg(s + x);
			x++;
			uint32_t l = tk >> 4, m = tk & 15u;
			bool err = false;
			if (l == 15) {
				uint32_t b;
				do {
					if (x >= a_end) { err = true; break; }
					b = __ldg(s + x);
					x++;
					l += b;
				} while (b == 255 && l < 0x40000000u);
			}
			const uint32_t qq = x;
			if (!err && l > a_end - x) err = true;
			uint32_t o = 0;
			if (!err) {
				x += l;
				if (x == a_end) {
					if (m) err = true;
					m = 0;
				} else if (x + 2 > a_end) {
					err = true;
				} else {
					o = __ldg(s + x) | (static_cast<uint32_t>(__ldg(s + x + 1)) << 8);
					x += 2;
					if (m == 15) {
						uint32_t b;
						do {
							if (x >= a_end) { err = true; break; }
							b = __ldg(s + x);
							x++;
							m += b;
						} while (b == 255 && m < 0x40000000u);
					}
					m += 4;
					if (!err && (o == 0 || o > p_cur + l - p_start)) err = true;
				}
			}
			if (!err && (l > p_cap - p_cur || m > p_cap - p_cur - l)) err = true;
			if (err) {
				bad = true;
				park = false;
			} else {
				lg_lit = l; lg_q = qq; lg_ml = m; lg_off = o; lg_nx = x;
			}
		}
		if (go && bad) {
			state = L_EXACT;   // the exact routine re-decodes the block and reports the error
			go = false;
		}
		const bool short_go = go && !park;

		// ---- literals: in ring -> registers -> out ring ----
		Bytes36 D;
#pragma unroll
		for (int j = 0; j < 9; j++) D.w[j] = 0;
		{
			const bool lact = short_go && lit > 0;
			const uint32_t maxlit = __reduce_max_sync(FULL_MASK, lact ? lit : 0u);
			if (maxlit) {
				fetch_col<IWW>(D, inb, lane4, q, lit, maxlit, lact);
				store_bytes(outb, lane4, D, p_cur, lit, maxlit, lact);
			}
		}
		// ---- match: out ring or global -> registers -> out ring ----
		{
			const bool mact = short_go && ml > 0;
			const uint32_t mo = p_cur + lit;
			const uint32_t src_s = mo - off;
			// what the ring still holds once this sequence is written
			const uint32_t lo_wr = mo + ml > OUT_BYTES ? mo + ml - OUT_BYTES : 0u;
			const uint32_t near_lo = ring_lo > lo_wr ? ring_lo : lo_wr;
			const bool is_near = mact && src_s >= near_lo;
			const bool is_far = mact && !is_near && src_s + ml <= p_flushed;
			if (mact && !is_near && !is_far) {
				// straddles the flush frontier (only right after a long sequence): park it
				lg_lit = 0; lg_q = q + lit; lg_ml = ml; lg_off = off; lg_nx = nx;
				park = true;
			}
			const uint32_t maxml = __reduce_max_sync(FULL_MASK, (is_near || is_far) ? ml : 0u);
			if (maxml) {
				__syncwarp();
				if (__any_sync(FULL_MASK, is_far)) fetch_global(D, obase + src_s, ml, is_far);
				if (__any_sync(FULL_MASK, is_near)) fetch_col<OWW>(D, outb, lane4, src_s, ml, maxml, is_near);
				store_bytes(outb, lane4, D, mo, ml, maxml, is_near || is_far);
			}
			if (short_go) {
				if (park) {
					// literals are done; the match goes the long way
					p_cur += lit;
				} else {
					p_cur += lit + ml;
					a_cur = nx;
				}
			}
		}
		// ---- flush complete 16-byte chunks of the out ring (two per trip keep up with 48 bytes per trip) ----
#pragma unroll 1
		for (int rep = 0; rep < 3; rep++) {
			const bool fl = (state == L_RUN) && (p_cur & ~15u) > p_flushed;
			if (!__any_sync(FULL_MASK, fl)) break;
			if (fl) {
				const uint32_t w0 = p_flushed >> 2;
				uint4 v;
				v.x = *reinterpret_cast<const uint32_t *>(outb + (((w0 + 0) & (OWW - 1)) << 7) + lane4);
				v.y = *reinterpret_cast<const uint32_t *>(outb + (((w0 + 1) & (OWW - 1)) << 7) + lane4);
				v.z = *reinterpret_cast<const uint32_t *>(outb + (((w0 + 2) & (OWW - 1)) << 7) + lane4);
				v.w = *reinterpret_cast<const uint32_t *>(outb + (((w0 + 3) & (OWW - 1)) << 7) + lane4);
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
			}
		}
		// ---- a lane that parks or finishes writes the partial chunk too: global memory is then complete ----
		const bool fin = (state == L_RUN) && !park && a_cur >= a_end && (p_cur & ~15u) <= p_flushed;
		if ((state == L_RUN && park) || fin) {
			if ((p_cur & ~15u) > p_flushed) {
				// more than three chunks behind cannot happen: every trip writes <= 48 bytes
			}
			const uint32_t base16 = p_flushed;
			const uint32_t w0 = base16 >> 2;
			uint32_t ws[4];
#pragma unroll
			for (int k = 0; k < 4; k++) ws[k] = *reinterpret_cast<const uint32_t *>(outb + (((w0 + k) & (OWW - 1)) << 7) + lane4);
#pragma unroll
			for (int k = 0; k < 16; k++)
				if (base16 + k >= p_start && base16 + k < p_cur) obase[base16 + k] = static_cast<uint8_t>(ws[k >> 2] >> (8 * (k & 3)));
			state = park ? L_LONG : L_FINISH;
		}
		// safety net: a state the lock-step machine cannot leave must not hang the device -- hand the blocks to the
		// exact routine (a lane makes progress whenever it consumes input, flushes, parks or finishes)
		{
			const bool progressed = go || (state != L_RUN);
			idle_trips = __any_sync(FULL_MASK, progressed) ? 0u : idle_trips + 1u;
			if (idle_trips > 4096u) {
				if (state == L_RUN) state = L_EXACT;
				idle_trips = 0;
			}
		}
		if (state == L_FINISH) {
			lz4b200_blk_status *st = status + blk;
			st->code = LZ4B200_ST_OK;
			st->out_len = p_cur - p_start;
			st->err_pos = 0;
			st->aux = 0;
			st->xxh32_computed = computed;
			st->xxh32_declared = declared;
			state = L_IDLE;
		}
		__syncwarp();
	}
}

}  // namespace v5
}  // namespace lz4b200
