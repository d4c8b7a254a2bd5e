// kernels_k6.cuh -- K6: the copy side of a chain (a linked frame, a 1 - 8 MiB block, an exact-placement retry) as ONE
// warp working in rounds on a shared-memory window, instead of K4's seven copier warps ordered by gates.
//
// A stream of ~10^5 sequences per MiB is serial twice: in its parse (a token's position follows from the previous
// literal length) and in its copies (a match reads what earlier matches wrote).  K4 broke the first chain with a
// speculative parallel parse and gave every batch of 32 sequences to a copier warp -- but the batches still had to be
// handed over in order, ~4 000 cycles per batch, 100 MB/s per stream.  Measured on the text corpus, though, the copy
// dependencies are shallow: a match depends on a match of the same 32 sequences ~5 levels deep at most (the longest
// chain through a 64 KiB block is ~800 levels for 6 300 sequences).  So:
//
//   window   the stream's last 64 KiB of output live in shared memory (a ring indexed by the output position, phase-
//            aligned with the global address): every match source is a shared-memory read, ~30 cycles, never L2
//   parse    a second warp, in steps of 1 KiB of compressed bytes staged in shared memory: for EVERY byte position "where
//            would the next token be if one started here", five rounds of pointer doubling over that table, then the
//            true token chain is read off in 32-token strides and filled in level by level -- no speculation, no
//            serial walk; descriptors carry output position, lengths and offset, so the copier reads nothing but
//            shared memory
//   batch    lane i takes sequence i of the batch the parser warp published: literals first (no dependencies), then
//            matches in ROUNDS -- a lane copies as soon as every earlier sequence of the batch whose output its source
//            touches is done (the rest of its past was final before the batch began); one ballot per round
//   copies   <= 16 bytes at a time per lane: five aligned words in, funnel shifts, then byte stores for the ragged
//            head and tail and word stores between (neighbouring lanes own neighbouring bytes of a word); periods
//            below 16 double their distance from piece to piece (pattern replication, lib/lz4ada.adb:893-903)
//   flush    after every batch the warp writes the complete 16-byte chunks of the window to global memory, coalesced
//   giants   a batch with a run of a kilobyte or more (zero pages, incompressible stretches) is copied in global
//            memory by the v2 batch routine (warp-wide copies, pattern replication) and the window is re-seeded
//
// Everything unusual (offset 0, a match reaching before the frame, capacity) makes the batch fail; the chain kernel then
// hands the rest of the chain to the exact routine (process_block: lib/lz4ada.adb:716-904 semantics), as K4 does.
#pragma once

#include "kernels_v2.cuh"

namespace lz4b200 {
namespace k6 {

constexpr uint32_t WIN = 65536;        // window bytes (a power of two: ring index = position & (WIN - 1))
constexpr uint32_t GIANT = 1024;       // runs this long, and ...
constexpr uint32_t BATCH_MAX = 8192;   // ... batches producing more than this, go through global memory

__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, const uint4 &v)
{
	asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// The copier's view of one chain.  Positions are chain-relative output positions; y = position + ph is the same in
// "aligned coordinates" (out_a[y] is the byte, out_a is 16-byte aligned), and the window keeps byte y at ring[y & 65535].
struct Window {
	uint32_t ring;        // shared address of the window
	uint8_t *out_a;       // chain output base, aligned down to 16 bytes
	uint32_t ph;          // misalignment of the chain output base
	uint64_t flushed;     // global memory is complete below this chain position
	uint64_t valid_from;  // the window holds the chain's bytes from here (up to what has been produced)
};

// n <= 16 bytes D (byte 0 = low byte of D0) into the window at aligned coordinate y.
__device__ __forceinline__ void put_piece(uint32_t ring, uint32_t y, uint32_t D0, uint32_t D1, uint32_t D2, uint32_t D3, uint32_t n)
{
	constexpr uint32_t WM = WIN - 1;
	const uint32_t b = y & 3u;
	uint32_t h = (4u - b) & 3u;   // bytes up to the next word boundary
	h = h < n ? h : n;
	if (h > 0) sts8(ring + (y & WM), D0);
	if (h > 1) sts8(ring + ((y + 1) & WM), D0 >> 8);
	if (h > 2) sts8(ring + ((y + 2) & WM), D0 >> 16);
	const uint32_t rest = n - h, s8 = h * 8u;
	const uint32_t E0 = __funnelshift_r(D0, D1, s8), E1 = __funnelshift_r(D1, D2, s8), E2 = __funnelshift_r(D2, D3, s8),
		       E3 = __funnelshift_r(D3, 0u, s8);
	const uint32_t yw = y + h, nw = rest >> 2;
	if (nw > 0) sts32(ring + (yw & WM), E0);
	if (nw > 1) sts32(ring + ((yw + 4) & WM), E1);
	if (nw > 2) sts32(ring + ((yw + 8) & WM), E2);
	if (nw > 3) sts32(ring + ((yw + 12) & WM), E3);
	const uint32_t t = rest & 3u, yt = yw + 4u * nw;
	const uint32_t Et = nw == 0 ? E0 : nw == 1 ? E1 : nw == 2 ? E2 : E3;
	if (t > 0) sts8(ring + (yt & WM), Et);
	if (t > 1) sts8(ring + ((yt + 1) & WM), Et >> 8);
	if (t > 2) sts8(ring + ((yt + 2) & WM), Et >> 16);
}

// Complete 16-byte chunks of the window below chain position `upto` go to global memory (the whole warp, coalesced).
__device__ __forceinline__ void flush_chunks(Window &w, uint64_t upto, int lane)
{
	constexpr uint32_t WM = WIN - 1;
	if (w.flushed == 0 && w.ph != 0) {
		// the chain starts in the middle of a 16-byte granule: its first 16 - ph bytes go out bytewise, once
		const uint32_t first = 16u - w.ph;
		if (upto < first) return;
		if (static_cast<uint32_t>(lane) < first) {
			const uint32_t y = w.ph + lane;
			w.out_a[y] = static_cast<uint8_t>(lds32(w.ring + (y & WM & ~3u)) >> (8u * (y & 3u)));
		}
		w.flushed = first;
	}
	const uint64_t y0 = w.flushed + w.ph, y1 = (upto + w.ph) & ~15ull;   // aligned coordinates, multiples of 16
	for (uint64_t y = y0 + 16ull * lane; y < y1; y += 512)
		*reinterpret_cast<uint4 *>(w.out_a + y) = lds128(w.ring + (static_cast<uint32_t>(y) & WM));
	if (y1 > y0) w.flushed = y1 - w.ph;
}

// ... and the ragged rest below `upto` bytewise (end of a chain, or before the window is abandoned).
__device__ __forceinline__ void flush_tail(Window &w, uint64_t upto, int lane)
{
	constexpr uint32_t WM = WIN - 1;
	flush_chunks(w, upto, lane);
	for (uint64_t c = w.flushed + lane; c < upto; c += 32) {
		const uint32_t y = static_cast<uint32_t>(c) + w.ph;
		w.out_a[c + w.ph] = static_cast<uint8_t>(lds32(w.ring + (y & WM & ~3u)) >> (8u * (y & 3u)));
	}
	__syncwarp();
}

// Fill the window with the chain's last bytes below `upto` out of global memory (after a batch that went through
// global memory).  Bytes [flushed, upto) stay the window's to flush: they are in global memory already and are
// simply written again with their chunk.
__device__ __forceinline__ void reseed(Window &w, uint64_t upto, int lane)
{
	constexpr uint32_t WM = WIN - 1;
	const uint64_t from = upto > WIN - 64 ? upto - (WIN - 64) : 0;
	const uint64_t y0 = (from + w.ph) & ~15ull, y1 = (upto + w.ph + 15) & ~15ull;
	for (uint64_t y = y0 + 16ull * lane; y < y1; y += 512)
		sts128(w.ring + (static_cast<uint32_t>(y) & WM), *reinterpret_cast<const uint4 *>(w.out_a + y));
	w.valid_from = y0 > w.ph ? y0 - w.ph : 0;
	const uint64_t fl = (upto + w.ph) & ~15ull;
	w.flushed = fl > w.ph ? fl - w.ph : 0;
	if (w.flushed == 0 && w.ph != 0 && upto >= 16u - w.ph) w.flushed = 16u - w.ph;   // (the first granule went out with the batch)
	__syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Parser <-> copier interface
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t CH = 1024;       // compressed bytes whose tokens one parse step enumerates
constexpr uint32_t RIN = 8192;      // staged compressed bytes (ring indexed by the aligned input coordinate)
constexpr uint32_t AHEAD = 3072;    // ... kept this far ahead of the parse position
constexpr uint32_t SLOTS = 32;      // batches in flight between parser and copier
constexpr uint32_t TOKMAX = CH / 3 + 32;
// next-table values >= CH end the walk of a step: the exact exit position, or FINAL (the block ends with this
// sequence), COMPLEX (a token for the serial routine), BEYOND (no token can start here: behind the block)
constexpr uint32_t FINAL = 0xfffdu, COMPLEX = 0xfffeu, BEYOND = 0xffffu;

struct Desc {             // 16 bytes per sequence
	uint32_t out_pos;     // frame-relative position of the first literal byte
	uint32_t lens;        // literal length | match length << 16 (0 = final literal-only sequence)
	uint32_t off;         // match offset
	uint32_t lit_pos;     // block-relative position of the first literal byte
};

enum : uint32_t { BK_SEQ = 0, BK_STORED = 1 };

struct Shared {
	Desc sd[SLOTS][32];
	uint32_t count[SLOTS], kind[SLOTS], ostart[SLOTS], total[SLOTS], cap_abs[SLOTS];
	uint32_t fb_lo[SLOTS], fb_hi[SLOTS], src_lo[SLOTS], src_hi[SLOTS], blk[SLOTS], in_pos[SLOTS], in_mis[SLOTS];
	uint16_t nx[6][CH];
	uint16_t tok[TOKMAX + 32];
	uint32_t produced, done_upto, fail, end_batch, fail_block;
	uint32_t copier_in;   // aligned input coordinate below which the copier needs no staged bytes any more
	uint32_t copier_blk;  // ... of this block (chain-relative index); earlier blocks are done
	__align__(16) uint8_t inr[RIN];
};

constexpr uint32_t SPIN_MAX = 1u << 22;   // polls of ~20 ns: a wait this long means a bug, not work -- give up to the exact routine
__device__ __forceinline__ uint32_t vld(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void vst(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// byte at aligned input coordinate xa out of the staged ring
__device__ __forceinline__ uint32_t in_byte(uint32_t inr, uint32_t xa)
{
	return (lds32(inr + (xa & (RIN - 1) & ~3u)) >> (8u * (xa & 3u))) & 255u;
}

// Where the next token starts if one starts at block position pos (< n).  Returns the block-relative position, or
// n for the final literal-only sequence; COMPLEX_POS for anything the table does not take (long extensions, bytes
// that are not staged, structural errors: the serial routine looks at those when the true chain reaches them).
constexpr uint32_t COMPLEX_POS = 0xffffffffu;
__device__ __forceinline__ uint32_t next_token(uint32_t inr, uint32_t mis, uint32_t pos, uint32_t n, uint32_t staged_hi,
					       uint32_t &lit, uint32_t &ml, uint32_t &lit_pos)
{
	const uint32_t tk = in_byte(inr, pos + mis);
	lit = tk >> 4;
	ml = tk & 15u;
	uint32_t p = pos + 1;
	if (lit == 15u) {
		if (p >= n || p >= staged_hi) return COMPLEX_POS;
		const uint32_t e = in_byte(inr, p + mis);
		if (e == 255u) return COMPLEX_POS;
		lit += e;
		p++;
	}
	lit_pos = p;
	const uint32_t q = p + lit;
	if (q > n) return COMPLEX_POS;
	if (q == n) {
		if (ml) return COMPLEX_POS;
		return n;
	}
	if (q + 2 > n || q + 2 > staged_hi) return COMPLEX_POS;
	uint32_t nxt = q + 2;
	if (ml == 15u) {
		if (nxt >= n || nxt >= staged_hi) return COMPLEX_POS;
		const uint32_t e = in_byte(inr, nxt + mis);
		if (e == 255u) return COMPLEX_POS;
		ml += e;
		nxt++;
	}
	ml += 4;
	return nxt;
}

// Stage the 16-byte granules [lo, hi) (aligned input coordinates, multiples of 16) of the block at sa into the ring.
__device__ __forceinline__ void stage(uint32_t inr, const uint8_t *sa, uint32_t lo, uint32_t hi, int lane)
{
	for (uint32_t g = lo + 16u * lane; g < hi; g += 512u) {
		const uint32_t dst = inr + (g & (RIN - 1));
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(sa + g) : "memory");
	}
}

// The parser warp: publishes batches of <= 32 sequence descriptors (or one stored block) per slot.
__device__ __forceinline__ void parser_role(Shared &sh, const lz4b200_chain &ch, const uint8_t *__restrict__ src,
					    const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, int lane)
{
	const uint32_t inr = static_cast<uint32_t>(__cvta_generic_to_shared(sh.inr));
	uint64_t pos = 0, frame_start = 0;   // chain-relative
	uint32_t ring = 0;                   // Output_Pos of the reference's Buffer (LZ4B200_BLK_RING_CAP blocks)
	uint32_t k = 0;                      // batches published
	bool stop = false;
	auto wait_slot = [&](uint32_t need) {   // room for `need` more batches; false = the copier gave up
		if (lane == 0) {
			uint32_t spins = 0;
			while (k + need - vld(&sh.done_upto) > SLOTS && vld(&sh.fail) == 0) {
				;
				if (++spins > SPIN_MAX) vst(&sh.fail, 1);
			}
		}
		__syncwarp();
		return vld(&sh.fail) == 0;
	};
	// staged bytes below this aligned coordinate of block i are the parser's to overwrite: the copier works on a later
	// batch of the block (or has finished everything published so far)
	auto released = [&](uint32_t i) -> uint32_t {
		if (vld(&sh.done_upto) == k) return 0xffff0000u;
		return vld(&sh.copier_blk) == i ? (vld(&sh.copier_in) & ~15u) : 0u;
	};
	for (uint32_t i = 0; i < ch.n_blocks && !stop; i++) {
		const uint32_t b = ch.first_block + i;
		const lz4b200_blk_desc d = desc[b];
		if (d.flags & LZ4B200_BLK_FIRST_OF_FRAME) { frame_start = pos; ring = 0; }
		const uint64_t fpos0 = pos - frame_start;
		const uint64_t room = ch.dst_cap - pos;
		uint32_t blk_cap = d.dst_cap;
		if (d.flags & LZ4B200_BLK_RING_CAP) {   // lib/lz4ada.adb:678-680: what is left of the caller's Buffer behind the ring cursor
			if (ring >= 65536u) ring = 0;
			blk_cap = d.dst_cap > ring ? d.dst_cap - ring : 0u;
		}
		const uint32_t cap = room < blk_cap ? static_cast<uint32_t>(room) : blk_cap;
		const uint8_t *s = src + d.src_off;
		const uint32_t n = d.src_len;
		const bool stored = (d.flags & LZ4B200_BLK_STORED) != 0;
		const bool ordinary = !(d.flags & LZ4B200_BLK_HASH_ONLY) && fpos0 + cap < 0xfff00000ull && !(stored && n > cap) && n < 0x7ff00000u;
		uint32_t computed = 0, declared = 0;
		bool okay = ordinary;
		if (okay && (d.flags & LZ4B200_BLK_HAS_CHECKSUM)) {   // Check_Checksum before any decoding, :672-676
			const uint8_t *t = s + n;
			declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
			computed = quad_xxh32_prologue(s, n, lane);
			okay = computed == declared;
		}
		uint32_t fpos = static_cast<uint32_t>(fpos0);
		auto meta = [&](uint32_t sl, uint32_t kind, uint32_t cnt, uint32_t ostart, uint32_t total, uint32_t in_pos, uint32_t mis) {
			sh.count[sl] = cnt;
			sh.kind[sl] = kind;
			sh.ostart[sl] = ostart;
			sh.total[sl] = total;
			sh.cap_abs[sl] = static_cast<uint32_t>(fpos0) + cap;
			sh.fb_lo[sl] = static_cast<uint32_t>(frame_start);
			sh.fb_hi[sl] = static_cast<uint32_t>(frame_start >> 32);
			sh.src_lo[sl] = static_cast<uint32_t>(d.src_off);
			sh.src_hi[sl] = static_cast<uint32_t>(d.src_off >> 32);
			sh.blk[sl] = i;
			sh.in_pos[sl] = in_pos;
			sh.in_mis[sl] = mis;
		};
		if (okay && stored) {
			// stored block (lib/lz4ada.adb:685-695): the copier owns the output window, so it does the copy
			if (!wait_slot(1)) { okay = false; }
			else {
				if (lane == 0) {
					meta(k % SLOTS, BK_STORED, 0, fpos, n, 0, 0);
					__threadfence_block();
					vst(&sh.produced, k + 1);
				}
				__syncwarp();
				k++;
				fpos += n;
			}
		} else if (okay) {
			const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(s) & 15u);
			const uint8_t *sa = s - mis;                       // aligned coordinate xa = block position + mis
			const uint32_t na16 = (n + mis + 15u) & ~15u;      // end of the block's granules
			uint32_t st_hi = 0;                                // granules below this are staged (requested)
			uint32_t ip = 0;
			// the copier may still be reading staged bytes of the previous block: wait until it has left it
			if (lane == 0) {
				uint32_t spins = 0;
				while (k != vld(&sh.done_upto) && vld(&sh.fail) == 0) {
					;
					if (++spins > SPIN_MAX) vst(&sh.fail, 1);
				}
			}
			__syncwarp();
			if (vld(&sh.fail) != 0) okay = false;
			while (ip < n && okay) {
				// ---- staging: AHEAD bytes beyond the parse position, never over bytes the copier still needs ----
				{
					uint32_t need = ((ip + mis + CH + 48u) + 15u) & ~15u;   // this step reads [ip, ip + CH + 32)
					need = need < na16 ? need : na16;
					const bool cold = need > st_hi;   // not requested by an earlier step (first step of a block, long jumps)
					// (never stage what lies behind the parse position: after a long literal run the gap may exceed the
					// ring, and two copies in flight to one slot land in no particular order)
					const uint32_t ip16 = (ip + mis) & ~15u;
					if (st_hi < ip16) st_hi = ip16;
					uint32_t want = ((ip + mis + CH + AHEAD) + 15u) & ~15u;
					want = want < na16 ? want : na16;
					const uint32_t rel = released(i);
					const uint32_t limit = rel > 0xffff0000u - RIN ? 0xffff0000u : rel + RIN;
					want = want < limit ? want : limit;
					if (want > st_hi) {
						stage(inr, sa, st_hi, want, lane);
						st_hi = want;
					}
					asm volatile("cp.async.commit_group;" ::: "memory");
					if (need > st_hi) {
						// the copier is far behind: wait for it to release ring space, then stage the rest
						if (lane == 0) {
							uint32_t spins = 0;
							for (;;) {
								const uint32_t rel2 = released(i);
								if (rel2 >= 0xffff0000u - RIN || rel2 + RIN >= need || vld(&sh.fail) != 0) break;
								;
								if (++spins > SPIN_MAX) vst(&sh.fail, 1);
							}
						}
						__syncwarp();
						if (vld(&sh.fail) != 0) { okay = false; break; }
						stage(inr, sa, st_hi, need, lane);
						st_hi = need;
						asm volatile("cp.async.commit_group;" ::: "memory");
					}
					// what this step reads was requested steps ago (only the newest group may still be in flight) -- unless cold
					if (cold) asm volatile("cp.async.wait_group 0;" ::: "memory");
					else asm volatile("cp.async.wait_group 1;" ::: "memory");
					__syncwarp();
				}
				const uint32_t staged_hi = st_hi - mis;   // block positions below this are in the ring
				// ---- (a) for every byte position of the step: where would the next token start? ----
				// (eight positions per lane at a time: the loads of all eight are in flight together)
#pragma unroll 1
				for (uint32_t x0 = lane; x0 < CH; x0 += 256) {
					uint32_t tk[8];
#pragma unroll
					for (int u = 0; u < 8; u++) tk[u] = in_byte(inr, ip + x0 + 32u * u + mis);
#pragma unroll
					for (int u = 0; u < 8; u++) {
						const uint32_t x = x0 + 32u * u, p0 = ip + x;
						uint32_t v = BEYOND;
						if (p0 < n) {
							const uint32_t l4 = tk[u] >> 4, m4 = tk[u] & 15u;
							if (l4 != 15u && m4 != 15u) {
								// the plain token: literals, offset, no extension bytes
								const uint32_t q = p0 + 1 + l4;
								v = q == n ? (m4 ? COMPLEX : FINAL) : q + 2 > n || q + 2 > staged_hi ? COMPLEX : q + 2 == n ? FINAL : q + 2 - ip;
							} else {
								uint32_t lit, ml, lp;
								const uint32_t nxt = next_token(inr, mis, p0, n, staged_hi, lit, ml, lp);
								v = nxt == COMPLEX_POS || lit > 60000u || nxt - ip >= 0xfff0u ? COMPLEX : nxt == n ? FINAL : nxt - ip;
							}
						}
						sh.nx[0][x] = static_cast<uint16_t>(v);
					}
				}
				__syncwarp();
				// ---- (b) pointer doubling: nx[j][x] = 2^j tokens on (values >= CH are exits and stay) ----
#pragma unroll
				for (int j = 1; j <= 5; j++) {
#pragma unroll 1
					for (uint32_t x0 = lane; x0 < CH; x0 += 256) {
						uint32_t v[8], w2[8];
#pragma unroll
						for (int u = 0; u < 8; u++) v[u] = sh.nx[j - 1][x0 + 32u * u];
#pragma unroll
						for (int u = 0; u < 8; u++) w2[u] = sh.nx[j - 1][v[u] < CH ? v[u] : 0u];
#pragma unroll
						for (int u = 0; u < 8; u++) sh.nx[j][x0 + 32u * u] = static_cast<uint16_t>(v[u] < CH ? w2[u] : v[u]);
					}
					__syncwarp();
				}
				// ---- (c) the true chain in strides of 32 tokens (every lane walks it: a dozen steps) ----
				uint32_t G = 0;
				{
					uint32_t t = 0;
					while (t < CH && G < TOKMAX / 32) {
						if (lane == 0) sh.tok[32 * G] = static_cast<uint16_t>(t);
						t = sh.nx[5][t];
						G++;
					}
				}
				__syncwarp();
				// ---- (d) fill in: stride 16, 8, 4, 2, 1 ----
#pragma unroll
				for (int j = 4; j >= 0; j--) {
					const uint32_t known = (G * 32u) >> (j + 1);
					for (uint32_t q = lane; q < known; q += 32) {
						const uint32_t t = sh.tok[q << (j + 1)];
						sh.tok[(q << (j + 1)) + (1u << j)] = t < CH ? sh.nx[j][t] : static_cast<uint16_t>(BEYOND);
					}
					__syncwarp();
				}
				// ---- (e) how many tokens start in this step, and where the chain leaves it ----
				uint32_t cnt = 0;
				for (uint32_t q = lane; q < G * 32u; q += 32) cnt += sh.tok[q] < CH ? 1u : 0u;
				cnt = __reduce_add_sync(FULL_MASK, cnt);
				const uint32_t last = sh.tok[cnt - 1];            // (cnt >= 1: position 0 is a token)
				const uint32_t exit_v = sh.nx[0][last];           // exact exit (>= CH), or COMPLEX
				const uint32_t n_plain = exit_v == COMPLEX ? cnt - 1 : cnt;   // tokens the fast path emits
				// ---- (f) descriptors, 32 tokens per batch ----
				const uint32_t nb = (n_plain + 31u) / 32u;
				if (nb) {
					if (!wait_slot(nb)) { okay = false; break; }
					for (uint32_t g = 0; g < nb; g++) {
						const uint32_t qi = g * 32u + lane;
						uint32_t lit = 0, ml = 0, lp = 0, off = 0;
						if (qi < n_plain) {
							const uint32_t p0 = ip + sh.tok[qi];
							next_token(inr, mis, p0, n, staged_hi, lit, ml, lp);
							if (ml) off = in_byte(inr, lp + lit + mis) | (in_byte(inr, lp + lit + 1 + mis) << 8);
						}
						const uint32_t len = lit + ml;
						uint32_t incl = len;
#pragma unroll
						for (int sft = 1; sft < 32; sft <<= 1) {
							const uint32_t v = __shfl_up_sync(FULL_MASK, incl, sft);
							if (lane >= sft) incl += v;
						}
						const uint32_t tot = __shfl_sync(FULL_MASK, incl, 31);
						const uint32_t sl = (k + g) % SLOTS;
						Desc dd;
						dd.out_pos = fpos + incl - len;
						dd.lens = lit | (ml << 16);
						dd.off = off;
						dd.lit_pos = lp;
						*reinterpret_cast<uint4 *>(&sh.sd[sl][lane]) = *reinterpret_cast<const uint4 *>(&dd);
						const uint32_t first_lp = __shfl_sync(FULL_MASK, lp, 0);
						if (lane == 0) {
							const uint32_t c = n_plain - g * 32u < 32u ? n_plain - g * 32u : 32u;
							meta(sl, BK_SEQ, c, fpos, tot, first_lp + mis, mis);
						}
						fpos += tot;
						if (tot > cap - (fpos - tot - static_cast<uint32_t>(fpos0))) okay = false;   // the exact path reports the overflow
					}
					__syncwarp();
					__threadfence_block();
					if (lane == 0) vst(&sh.produced, k + nb);
					k += nb;
					if (!okay) break;
				}
				if (exit_v != COMPLEX) {
					ip = exit_v == FINAL ? n : ip + exit_v;
					continue;
				}
				// ---- a token the table does not take (long length extensions ...): the serial routine, one sequence ----
				{
					const uint32_t p0 = ip + last;
					uint32_t lp = 0, lit = 0, ml = 0, nxt = 0;
					bool fine = true;
					if (lane == 0) fine = parse_token(s, n, p0, lp, lit, ml, nxt);
					fine = __shfl_sync(FULL_MASK, fine ? 1 : 0, 0) != 0;
					lp = __shfl_sync(FULL_MASK, lp, 0);
					lit = __shfl_sync(FULL_MASK, lit, 0);
					ml = __shfl_sync(FULL_MASK, ml, 0);
					nxt = __shfl_sync(FULL_MASK, nxt, 0);
					if (!fine || lit + ml > cap - (fpos - static_cast<uint32_t>(fpos0)) || !wait_slot(1)) { okay = false; break; }
					if (lane == 0) {
						const uint32_t sl = k % SLOTS;
						Desc dd;
						dd.out_pos = fpos;
						dd.lens = lit | (ml << 16);
						dd.off = ml ? (ld_u8<true>(s + lp + lit) | (ld_u8<true>(s + lp + lit + 1) << 8)) : 0u;
						dd.lit_pos = lp;
						sh.sd[sl][0] = dd;
						meta(sl, BK_SEQ, 1, fpos, lit + ml, lp + mis, mis);
						__threadfence_block();
						vst(&sh.produced, k + 1);
					}
					__syncwarp();
					k++;
					fpos += lit + ml;
					ip = nxt;
				}
			}
		}
		if (okay) {
			// provisional: stands unless the copier gives up on one of this block's batches
			if (lane == 0) {
				status[b].code = LZ4B200_ST_OK;
				status[b].out_len = fpos - static_cast<uint32_t>(fpos0);
				status[b].err_pos = 0;
				status[b].aux = 0;
				status[b].xxh32_computed = computed;
				status[b].xxh32_declared = declared;
			}
			pos += fpos - static_cast<uint32_t>(fpos0);
			ring += fpos - static_cast<uint32_t>(fpos0);
		} else {
			// anything out of the ordinary: drain the pipeline, then the exact routine takes over from this block
			if (lane == 0) atomicMin(&sh.fail_block, i);
			stop = true;
		}
	}
	if (lane == 0) {
		__threadfence_block();
		vst(&sh.end_batch, k);
	}
}

// One batch of <= 32 sequences of the block whose payload starts at blk_src (staged in the ring `inr` at aligned
// coordinate = block position + mis).  C0 = chain position where the batch's output starts, ostart = the same relative
// to its frame (what a match may reach back), cap_abs = frame-relative capacity.  Returns false when the batch needs the
// exact path.  Warp-uniform result.
__device__ __forceinline__ bool copy_batch_rounds(Window &w, const uint8_t *__restrict__ blk_src, uint32_t inr, uint32_t mis,
						  uint8_t *frame_out, uint64_t C0, uint32_t ostart, uint32_t total, uint32_t cap_abs,
						  const Desc *sd, uint32_t cnt, int lane, uint8_t *tile)
{
	constexpr uint32_t WM = WIN - 1;
	const bool act = static_cast<uint32_t>(lane) < cnt;
	uint32_t lit_pos = 0, lit = 0, ml = 0, off = 0, r = total;
	if (act) {
		const uint4 raw = *reinterpret_cast<const uint4 *>(sd + lane);
		r = raw.x - ostart;   // batch-relative start of this sequence's output
		lit = raw.y & 0xffffu;
		ml = raw.y >> 16;
		off = raw.z;
		lit_pos = raw.w;
	}
	const uint32_t rm = r + lit;          // batch-relative start of the match
	const uint32_t nxt = rm + ml;         // ... and of the next sequence
	const bool bad = ml && (off == 0 || off > ostart + rm);   // lib/lz4ada.adb:766-772, :864-874
	if (__any_sync(FULL_MASK, bad) || total > cap_abs - ostart) return false;

	if (total > BATCH_MAX || __any_sync(FULL_MASK, lit >= GIANT || ml >= GIANT)) {
		// ---- long runs: the v2 batch routine in global memory (warp-wide copies), then a fresh window ----
		flush_tail(w, C0, lane);
		SeqDesc *old = reinterpret_cast<SeqDesc *>(tile + TILE_BYTES + 32);
		if (act) *reinterpret_cast<uint2 *>(old + lane) = make_uint2(lit_pos, lit | (ml << 16));
		__syncwarp();
		uint32_t t2 = 0;
		if (!copy_batch<false>(blk_src, frame_out, ostart, cap_abs, old, cnt, lane, tile, t2)) return false;
		__syncwarp();
		reseed(w, C0 + total, lane);
		return true;
	}

	const uint32_t Y0 = static_cast<uint32_t>(C0) + w.ph;   // aligned coordinate of the batch start (ring index = low bits)
	// ---- literals: no dependencies (Write_Output, :790-824); out of the staged ring while the batch's span fits it ----
	{
		const uint32_t first_in = __shfl_sync(FULL_MASK, lit_pos, 0);
		const uint32_t span = __reduce_max_sync(FULL_MASK, act ? lit_pos + lit - first_in : 0u);
		const bool staged = span <= RIN - 64;
		uint32_t done = 0;
		while (__any_sync(FULL_MASK, done < lit)) {
			if (done < lit) {
				const uint32_t n = lit - done < 16u ? lit - done : 16u;
				uint32_t x0, x1, x2, x3, x4, bs;
				if (staged) {
					const uint32_t xa = lit_pos + done + mis;
					const uint32_t wa = xa & (RIN - 1) & ~3u;
					x0 = lds32(inr + wa);
					x1 = lds32(inr + ((wa + 4) & (RIN - 1)));
					x2 = lds32(inr + ((wa + 8) & (RIN - 1)));
					x3 = lds32(inr + ((wa + 12) & (RIN - 1)));
					x4 = lds32(inr + ((wa + 16) & (RIN - 1)));
					bs = (xa & 3u) * 8u;
				} else {
					const uint8_t *q = blk_src + lit_pos + done;
					const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(q) & 3u);
					const uint32_t *qa = reinterpret_cast<const uint32_t *>(q - m);
					x0 = __ldg(qa); x1 = __ldg(qa + 1); x2 = __ldg(qa + 2); x3 = __ldg(qa + 3); x4 = __ldg(qa + 4);
					bs = m * 8u;
				}
				put_piece(w.ring, Y0 + r + done, __funnelshift_r(x0, x1, bs), __funnelshift_r(x1, x2, bs), __funnelshift_r(x2, x3, bs),
					  __funnelshift_r(x3, x4, bs), n);
				done += n;
			}
		}
	}
	// ---- which earlier matches of the batch does my source touch?  Outputs are in lane order, so it is a range ----
	const int32_t ss = static_cast<int32_t>(rm) - static_cast<int32_t>(off);       // batch-relative source start (< 0: before the batch)
	const int32_t se = ss + static_cast<int32_t>(ml < off ? ml : off);              // ... and end (self-overlap: up to my own start)
	uint32_t dep = 0;
	if (__any_sync(FULL_MASK, ml != 0 && se > 0)) {
		int lo = 0, hi = 0;
#pragma unroll
		for (int step = 16; step; step >>= 1) {
			const int32_t a = static_cast<int32_t>(__shfl_sync(FULL_MASK, nxt, lo + step - 1));
			const int32_t b = static_cast<int32_t>(__shfl_sync(FULL_MASK, rm, hi + step - 1));
			if (a <= ss) lo += step;    // sequences that end at or before my source start
			if (b < se) hi += step;     // sequences whose match starts before my source end
		}
		// lo = first sequence reaching into my source, hi = one past the last
		if (ml && se > 0 && hi > lo) {
			const uint32_t upto = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
			dep = upto & ~((1u << lo) - 1u) & ((1u << lane) - 1u);
		}
	}
	__syncwarp();
	// ---- matches in rounds (Output_With_History, :845-904) ----
	const uint64_t E = C0 + total;
	uint32_t pending = __ballot_sync(FULL_MASK, ml != 0);
	dep &= pending;
	bool mine = ml != 0;
	while (pending) {
		const bool ready = mine && (dep & pending) == 0;
		if (ready) {
			uint32_t rem = ml, d = off, y = Y0 + rm;
			// a source the window no longer holds (more than 64 KiB before the end of this batch, or from before the
			// window was seeded): global memory has it -- it is far below the flush frontier
			const uint64_t S = C0 + rm - off;
			const bool old = S + WIN < E + 32 || S < w.valid_from;
			while (rem) {
				uint32_t n = rem < 16u ? rem : 16u;
				n = n < d ? n : d;
				uint32_t D0, D1, D2, D3;
				if (!old) {
					const uint32_t sy = y - d;
					const uint32_t wa = sy & WM & ~3u;
					const uint32_t x0 = lds32(w.ring + wa), x1 = lds32(w.ring + ((wa + 4) & WM)), x2 = lds32(w.ring + ((wa + 8) & WM)),
						       x3 = lds32(w.ring + ((wa + 12) & WM)), x4 = lds32(w.ring + ((wa + 16) & WM));
					const uint32_t bs = (sy & 3u) * 8u;
					D0 = __funnelshift_r(x0, x1, bs);
					D1 = __funnelshift_r(x1, x2, bs);
					D2 = __funnelshift_r(x2, x3, bs);
					D3 = __funnelshift_r(x3, x4, bs);
				} else {
					const uint8_t *g = w.out_a + (C0 + (y - Y0) - d + w.ph);
					uint32_t t[4] = {0, 0, 0, 0};
#pragma unroll
					for (int k = 0; k < 16; k++)
						if (static_cast<uint32_t>(k) < n) t[k >> 2] |= static_cast<uint32_t>(g[k]) << (8 * (k & 3));
					D0 = t[0]; D1 = t[1]; D2 = t[2]; D3 = t[3];
				}
				put_piece(w.ring, y, D0, D1, D2, D3, n);
				y += n;
				rem -= n;
				if (d < 16u && n == d) d <<= 1;   // the pattern has doubled
			}
			mine = false;
		}
		__syncwarp();
		pending &= ~__ballot_sync(FULL_MASK, ready);
	}
	flush_chunks(w, E, lane);
	__syncwarp();
	return true;
}

// The copier warp: batches in order.
__device__ __forceinline__ void copier_role(Shared &sh, uint32_t ring, uint8_t *tile, const lz4b200_chain &ch,
					    const uint8_t *__restrict__ src, uint8_t *out, int lane, uint32_t dbg)
{
	Window w;
	w.ring = ring;
	w.ph = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(out) & 15u);
	w.out_a = out - w.ph;
	w.flushed = 0;
	w.valid_from = 0;
	const uint32_t inr = static_cast<uint32_t>(__cvta_generic_to_shared(sh.inr));
	uint64_t produced_to = 0;   // chain position behind the last batch copied
	for (uint32_t k = 0;; k++) {
		uint32_t go = 0;
		if (lane == 0) {
			uint32_t spins = 0;
			for (;;) {
				if (vld(&sh.produced) > k) { go = 1; break; }
				if (vld(&sh.end_batch) <= k) { go = 0; break; }
				;
				if (++spins > 4u * SPIN_MAX) { vst(&sh.fail, 1); atomicMin(&sh.fail_block, 0u); go = 0; break; }
			}
		}
		go = __shfl_sync(FULL_MASK, go, 0);
		if (!go) break;
		__threadfence_block();
		const uint32_t sl = k % SLOTS;
		const uint32_t cnt = sh.count[sl], kind = sh.kind[sl], ostart = sh.ostart[sl], total = sh.total[sl], cap_abs = sh.cap_abs[sl];
		const uint32_t blk = sh.blk[sl];
		const uint64_t fbase = sh.fb_lo[sl] | (static_cast<uint64_t>(sh.fb_hi[sl]) << 32);
		const uint64_t soff = sh.src_lo[sl] | (static_cast<uint64_t>(sh.src_hi[sl]) << 32);
		if (lane == 0) {
			// what the parser may overwrite in the staged ring: everything below this batch's first literal
			vst(&sh.copier_in, sh.in_pos[sl]);
			vst(&sh.copier_blk, blk);
		}
		// batches from the first failing block on are redone by the exact routine: skip them
		const bool skip = blk >= vld(&sh.fail_block);
		bool okay = true;
		const uint64_t C0 = fbase + ostart;
		if (!skip && !(dbg & 1u)) {
			if (kind == BK_STORED) {
				// a stored block (lib/lz4ada.adb:685-695): straight through global memory, then a fresh window
				flush_tail(w, C0, lane);
				warp_copy<true>(out + C0, src + soff, total, lane);
				__syncwarp();
				reseed(w, C0 + total, lane);
			} else {
				okay = copy_batch_rounds(w, src + soff, inr, sh.in_mis[sl], out + fbase, C0, ostart, total, cap_abs, sh.sd[sl], cnt, lane, tile);
				// (timing aid, LZ4B200_K6_DBG = 2 r: every batch is copied 1 + r times -- the same bytes again)
				for (uint32_t rep = 0; rep < (dbg >> 1) && okay; rep++)
					copy_batch_rounds(w, src + soff, inr, sh.in_mis[sl], out + fbase, C0, ostart, total, cap_abs, sh.sd[sl], cnt, lane, tile);
			}
			if (okay) produced_to = C0 + total;
			else flush_tail(w, C0, lane);   // everything before the failing batch is final output
		}
		__syncwarp();
		if (lane == 0) {
			if (!okay) {
				atomicMin(&sh.fail_block, blk);   // where the exact routine restarts
				vst(&sh.fail, 1);
			}
			__threadfence_block();
			vst(&sh.done_upto, k + 1);
		}
		__syncwarp();
	}
	if (vld(&sh.fail) == 0) flush_tail(w, produced_to, lane);
	__threadfence_block();
}

}  // namespace k6
}  // namespace lz4b200
