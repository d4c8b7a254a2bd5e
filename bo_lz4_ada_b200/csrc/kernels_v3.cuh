// kernels_v3.cuh -- K1 third generation: one CTA per block, the block's whole output window and a
// window of its compressed bytes live in shared memory.
//
// What v2 (kernels_v2.cuh) measured on the 4 GiB text corpus: 35 warp instructions per LZ4 sequence,
// 45 GB of DRAM traffic for 6.2 GB of algorithmic bytes, because 16 000 blocks are open at once and
// every one keeps 64 KiB of output alive in L2 / DRAM for its matches (profiles/r01_final_k1_*).
// v3 keeps only two blocks open per SM and never touches global memory between the load of the
// compressed bytes and the final aligned flush of the finished block:
//
//   load    the next <= 33 KiB of the block's payload -> shared memory (cp.async, 16 bytes per lane)
//   parse   256 threads, one 132-byte segment each: walk the token chain from a *guessed* start, then
//           iterate "my entry = my left neighbour's exit" to the fixed point.  Chains started at a
//           wrong byte merge with the true chain within a few tokens, so two walks normally suffice;
//           the fixed point is exact whatever the data (worst case it degenerates to a serial walk).
//   emit    every thread re-walks its (now true) segment: token positions go to a compact table,
//           short literal runs are copied lane-per-sequence straight into the output window
//           (literals depend on nothing), batch start positions come from a block-wide scan.
//   match   batches of 32 sequences, round-robin over the 8 decode warps, lane per sequence.  A batch
//           may copy every match whose source lies below the in-order frontier at once; the rest
//           waits for the previous batch, then runs in dependency rounds (exact dependency masks).
//   flush   shared-memory window -> global, aligned 16-byte stores.
//
// A ninth warp hashes the compressed payloads of the CTA's blocks (eight serial XXH32 chains, one per
// quad: Check_Checksum, lib/lz4ada.adb:698-707) while the other eight decode.
//
// Anything unusual -- every error condition, matches reaching before the block, a single sequence
// larger than the window, blocks that may produce more than 64 KiB -- is handed to the exact routine
// (process_block) or to the v2 group decoder, which own the reference's error semantics
// (lib/lz4ada.adb:716-904).
#pragma once

#include "kernels_v2.cuh"

namespace lz4b200 {
namespace v3 {

constexpr int NW = 8;                        // decode warps
constexpr int NT = NW * 32;                  // decode threads = parse segments per window
constexpr int CTA_THREADS = NT + 32;         // + the hash warp
constexpr int MAX_NB = 8;                    // blocks per CTA (one per quad of the hash warp)
constexpr uint32_t SEG = 132;                // compressed bytes per parse segment: 33 words, so that
                                             // lane j's bytes sit in bank (j + k) mod 32
constexpr uint32_t WIN = NT * SEG;           // compressed window: 33 792 bytes
constexpr uint32_t CW_BYTES = WIN + 48;      // + alignment phase of the source address + rounding
constexpr uint32_t OUT_BYTES = 65536;        // output window = the largest block of the fast path
constexpr uint32_t NTOK = 6656;              // sequences per window (a 64 KiB text block has ~6 300)
constexpr uint32_t NBATCH = NTOK / 32;
constexpr uint32_t LIT_EMIT = 24;            // literal runs up to this length are copied by the emit walk

enum : uint32_t { W_OK = 0, W_CUT = 1, W_BAD = 2 };
// profile slots (cycles of thread 0 per phase; counts)
enum : int { P_LOAD = 0, P_PARSE, P_SCAN, P_EMIT, P_MATCH, P_FLUSH, P_EXACT, P_BLOCKS, P_WINDOWS, P_ITERS, P_FALLBACK,
	     P_BATCHES, P_ROUNDS, P_EARLY, P_GATEWAIT, P_COOP, PROF_N };

struct Ctl {
	uint32_t xs[NT];          // exit position of every segment (window-relative)
	uint32_t wc[NW], wo[NW];  // per-warp totals of the block-wide scan
	uint32_t fin_upto;        // batches of the current window that are final, in order
	uint32_t fail;            // the fast path gave up on the current block
	uint32_t T, O, wend;      // window summary: sequences, output bytes, compressed bytes consumed
	uint32_t pad[3];
	unsigned long long prof[PROF_N];   // LZ4B200_PROF=1: cycles per phase and event counts of this CTA
};

constexpr size_t SMEM_BYTES = OUT_BYTES + CW_BYTES + NTOK * 2 + (NBATCH + 4) * 4 + sizeof(Ctl);
static_assert(SMEM_BYTES <= 115712, "two CTAs per SM");
static_assert(CW_BYTES % 16 == 0 && (NTOK * 2) % 16 == 0, "alignment of the carved regions");

// Barriers among the decode threads only (the hash warp runs free).
__device__ __forceinline__ void dbar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
__device__ __forceinline__ bool dbar_or(bool p)
{
	uint32_t r;
	asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		     : "=r"(r) : "r"(static_cast<uint32_t>(p)), "n"(NT) : "memory");
	return r != 0;
}
__device__ __forceinline__ uint32_t dbar_popc(bool p)
{
	uint32_t r;
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.popc.u32 %0, 1, %2, q;\n\t}"
		     : "=r"(r) : "r"(static_cast<uint32_t>(p)), "n"(NT) : "memory");
	return r;
}

struct Walk { uint32_t x, c, o, st; };

// Walk the token chain from window position p until it leaves [.., seg_hi) (Decompress_Sequence,
// lib/lz4ada.adb:737-777, lengths by Process_Variable_Length :724-735).  cw = the window's bytes,
// wlen of them valid, `last` = the window reaches the end of the block.  A sequence that needs
// bytes beyond the window stops the walk in front of its token: W_CUT (the next window starts
// there) or, at the end of the block, W_BAD (truncated: the exact routine reports it).
// EMIT: record the tokens from index idx on and copy short literal runs to out + opos.
template <bool EMIT>
__device__ __forceinline__ Walk walk_segment(const uint8_t *cw, uint32_t p, uint32_t seg_hi, uint32_t wlen, bool last,
					     uint32_t idx, uint32_t opos, uint16_t *tokpos, uint32_t *bout, uint8_t *out)
{
	Walk r;
	r.c = 0; r.o = 0; r.st = W_OK;
	while (p < seg_hi) {
		const uint32_t tk = cw[p];
		uint32_t lit = tk >> 4, q = p + 1;
		bool cut = false;
		if (lit == 15) {
			uint32_t b;
			do {
				if (q >= wlen) { cut = true; break; }
				b = cw[q++];
				lit += b;
			} while (b == 255);
		}
		const uint32_t e = q + lit;   // end of the literals
		uint32_t ml = 0, nx = e;
		if (!cut) {
			if (e > wlen) {
				cut = true;
			} else if (last && e == wlen) {
				// final literal-only sequence (:752-764); a match nibble here is an error
				if (tk & 15) { r.st = W_BAD; break; }
			} else if (e + 2 > wlen) {
				cut = true;
			} else {
				nx = e + 2;
				ml = tk & 15;
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
		if (cut) { r.st = last ? W_BAD : W_CUT; break; }
		if (EMIT) {
			tokpos[idx] = static_cast<uint16_t>(p);
			if ((idx & 31u) == 0) bout[idx >> 5] = opos;
			if (lit <= LIT_EMIT) {
				for (uint32_t i = 0; i < lit; i++) out[opos + i] = cw[q + i];
			}
			idx++;
			opos += lit + ml;
		}
		r.c++;
		r.o += lit + ml;
		p = nx;
	}
	r.x = p;
	return r;
}

// dst[0..n) = src[0..n), both in shared memory, n bytes disjoint or dst - src >= n.
__device__ __forceinline__ void smem_copy(uint8_t *d, const uint8_t *s, uint32_t n, int lane)
{
	for (uint32_t i = lane; i < n; i += 32) d[i] = s[i];
}

// One match by the whole warp inside the output window, any length, any overlap
// (Output_With_History phases I and R, lib/lz4ada.adb:876-903).
__device__ __forceinline__ void coop_match(uint8_t *out, uint32_t mo, uint32_t off, uint32_t ml, int lane)
{
	uint8_t *d = out + mo;
	if (off >= ml) {
		smem_copy(d, d - off, ml, lane);
		return;
	}
	uint32_t done = 0;
	if (off < 32) {
		const uint32_t m0 = ml < 32 ? ml : 32;
		if (static_cast<uint32_t>(lane) < m0) d[lane] = *(d - off + (static_cast<uint32_t>(lane) % off));
		done = m0;
		__syncwarp();
	}
	while (done < ml) {
		// copy from a whole number of periods back: the source of every chunk is already valid
		const uint32_t avail = off + done;
		const uint32_t L = avail - (avail % off);
		const uint32_t chunk = (ml - done) < L ? (ml - done) : L;
		smem_copy(d + done, d + done - L, chunk, lane);
		done += chunk;
		__syncwarp();
	}
}

// Short non-overlapping match by one lane: <= 32 bytes through registers (three aligned 16-byte
// loads, a byte shift, byte stores).  maxml = warp-uniform bound on ml among the calling lanes.
__device__ __forceinline__ void copy_simple(uint8_t *out, uint32_t src_s, uint32_t mo, uint32_t ml, uint32_t maxml,
					    bool active)
{
	if (!active) return;
	const uint32_t m = src_s & 15u;
	const uint4 *base = reinterpret_cast<const uint4 *>(out + (src_s - m));
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
	uint8_t *dp = out + mo;
#pragma unroll
	for (int c8 = 0; c8 < 4; c8++) {
		if (static_cast<uint32_t>(c8 * 8) < maxml) {
#pragma unroll
			for (int k = 0; k < 8; k++)
				if (static_cast<uint32_t>(c8 * 8 + k) < ml) dp[c8 * 8 + k] = static_cast<uint8_t>(r[c8] >> (8 * k));
		}
	}
}

// One batch (sequences 32k .. 32k+31 of the window) by one warp.  Never leaves early: the batch is
// always published as final so that the warps behind it cannot wait forever; a sequence the fast
// path does not take (offset 0, match reaching before the block) raises ctl->fail instead.
__device__ __forceinline__ void match_batch(uint8_t *out, const uint8_t *cw, const uint16_t *tokpos, const uint32_t *bout,
					    uint32_t k, uint32_t T, uint32_t wlen, bool last, Ctl *ctl, int lane, bool prof)
{
	uint32_t n_rounds = 0, n_early = 0, n_coop = 0;
	long long t_gate = 0;
	const uint32_t idx = k * 32 + lane;
	const bool act = idx < T;
	const uint32_t B0 = bout[k];
	uint32_t lit = 0, ml = 0, off = 0, q = 0;
	if (act) {
		const uint32_t t = tokpos[idx];
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
	const uint32_t len = lit + ml;
	uint32_t incl = len;
#pragma unroll
	for (int s = 1; s < 32; s <<= 1) {
		const uint32_t v = __shfl_up_sync(FULL_MASK, incl, s);
		if (lane >= s) incl += v;
	}
	const uint32_t out_pos = B0 + incl - len;   // block-relative start of this sequence's output
	const uint32_t mo = out_pos + lit;          // ... and of its match
	const bool bad = ml && (off == 0 || off > mo);
	if (__any_sync(FULL_MASK, bad)) {
		if (lane == 0) ctl->fail = 1;
	}
	bool done = (ml == 0) || bad;

	// long literal runs (the emit walk skipped them): whole warp, one run at a time
	uint32_t big = __ballot_sync(FULL_MASK, act && lit > LIT_EMIT);
	while (big) {
		const int j = __ffs(big) - 1;
		big &= big - 1;
		smem_copy(out + __shfl_sync(FULL_MASK, out_pos, j), cw + __shfl_sync(FULL_MASK, q, j),
			  __shfl_sync(FULL_MASK, lit, j), lane);
	}

	const uint32_t src_s = mo - off;
	const uint32_t src_e = src_s + (ml < off ? ml : off);   // self-overlap: the source ends where the match starts
	const bool simple = ml <= 32 && off >= ml;

	// dep = earlier sequences of this batch whose output overlaps my source (binary search over the
	// monotone output positions; conservative: literal parts count too)
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

	// ---- before the gate: matches whose source is already below the in-order frontier ----
	volatile uint32_t *fin = &ctl->fin_upto;
	uint32_t f = 0;
	if (lane == 0) f = *fin;
	f = __shfl_sync(FULL_MASK, f, 0);
	if (f < k) {
		__threadfence_block();
		const uint32_t F = bout[f];   // every byte below the start of the oldest open batch is final
		const bool early = !done && simple && src_e <= F;
		const uint32_t em = __ballot_sync(FULL_MASK, early);
		if (__popc(em) >= 4) {
			const uint32_t maxml = __reduce_max_sync(FULL_MASK, early ? ml : 0u);
			copy_simple(out, src_s, mo, ml, maxml, early);
			done = done || early;
			n_early++;
		}
		if (prof) t_gate = clock64();
		if (lane == 0) {
			while (*fin < k) __nanosleep(20);
		}
		__syncwarp();
		if (prof) t_gate = clock64() - t_gate;
	}
	__threadfence_block();

	// ---- behind the gate: everything before B0 is final; rounds over the batch's own dependencies ----
	uint32_t undone = __ballot_sync(FULL_MASK, !done);
	while (undone) {
		const bool ready = !done && (dep & undone) == 0;
		const bool rs = ready && simple;
		if (__any_sync(FULL_MASK, rs)) {
			const uint32_t maxml = __reduce_max_sync(FULL_MASK, rs ? ml : 0u);
			copy_simple(out, src_s, mo, ml, maxml, rs);
		}
		uint32_t cx = __ballot_sync(FULL_MASK, ready && !simple);
		while (cx) {
			const int j = __ffs(cx) - 1;
			cx &= cx - 1;
			__syncwarp();
			coop_match(out, __shfl_sync(FULL_MASK, mo, j), __shfl_sync(FULL_MASK, off, j), __shfl_sync(FULL_MASK, ml, j), lane);
			n_coop++;
		}
		n_rounds++;
		done = done || ready;
		__syncwarp();
		undone = __ballot_sync(FULL_MASK, !done);
	}
	__syncwarp();
	__threadfence_block();
	if (lane == 0) *fin = k + 1;
	if (prof && lane == 0) {
		atomicAdd(&ctl->prof[P_BATCHES], 1ull);
		atomicAdd(&ctl->prof[P_ROUNDS], static_cast<unsigned long long>(n_rounds));
		atomicAdd(&ctl->prof[P_EARLY], static_cast<unsigned long long>(n_early));
		atomicAdd(&ctl->prof[P_COOP], static_cast<unsigned long long>(n_coop));
		atomicAdd(&ctl->prof[P_GATEWAIT], static_cast<unsigned long long>(t_gate));
	}
}

// Stored block (lib/lz4ada.adb:685-695) by the eight decode warps, global -> global.
__device__ __forceinline__ void cta_copy_stored(uint8_t *d, const uint8_t *s, uint32_t n, int warp, int lane)
{
	const uint32_t slice = ((n + NW - 1) / NW + 15u) & ~15u;
	const uint32_t lo = slice * warp;
	if (lo >= n) return;
	const uint32_t cnt = n - lo < slice ? n - lo : slice;
	warp_copy<true>(d + lo, s + lo, cnt, lane);
}

// The decode role: blocks first .. first + cnt - 1, one after the other.
__device__ __forceinline__ void decode_role(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks, uint32_t first,
					    uint32_t cnt, const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status,
					    uint8_t *smem, unsigned long long *gprof)
{
	const bool prof = gprof != nullptr;
	long long tp = 0;
	// thread 0 charges the cycles since the last mark to a phase
#define V3_MARK(slot)                                                   \
	do {                                                            \
		if (prof && threadIdx.x == 0) {                         \
			const long long now_ = clock64();               \
			ctl->prof[slot] += static_cast<unsigned long long>(now_ - tp); \
			tp = now_;                                      \
		}                                                       \
	} while (0)
#define V3_COUNT(slot, v)                                               \
	do {                                                            \
		if (prof && threadIdx.x == 0) ctl->prof[slot] += (v);   \
	} while (0)
	uint8_t *out = smem;
	uint8_t *cwbuf = out + OUT_BYTES;
	uint16_t *tokpos = reinterpret_cast<uint16_t *>(cwbuf + CW_BYTES);
	uint32_t *bout = reinterpret_cast<uint32_t *>(tokpos + NTOK);
	Ctl *ctl = reinterpret_cast<Ctl *>(bout + NBATCH + 4);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (prof) {
		if (tid < PROF_N) ctl->prof[tid] = 0;
		dbar();
		tp = clock64();
	}

	for (uint32_t bi = 0; bi < cnt; bi++) {
		const uint32_t b = first + bi;
		const lz4b200_blk_desc d = desc[b];
		if (d.flags & LZ4B200_BLK_NOT_K1) continue;
		const uint8_t *s = src + d.src_off;
		uint8_t *o = dst + d.dst_off;
		if (!(d.flags & LZ4B200_BLK_HASH_ONLY) && (d.flags & LZ4B200_BLK_STORED) && d.src_len <= d.dst_cap) {
			cta_copy_stored(o, s, d.src_len, warp, lane);
			if (tid == 0) {
				status[b].code = LZ4B200_ST_OK;
				status[b].out_len = d.src_len;
				status[b].err_pos = 0;
				status[b].aux = 0;
			}
			continue;
		}
		if ((d.flags & (LZ4B200_BLK_HASH_ONLY | LZ4B200_BLK_STORED)) || d.dst_cap > OUT_BYTES) {
			// not a block for this kernel's fast path: v2 group decoder (one block) on warp 0
			dbar();
			if (warp == 0)
				decode_group<1>(src, dst, n_blocks, b, desc, status, reinterpret_cast<SeqDesc *>(cwbuf),
						cwbuf + 512, lane);
			dbar();
			continue;
		}

		const uint32_t n = d.src_len, cap = d.dst_cap;
		uint32_t ip = 0, op = 0;
		bool fail = false;
		if (tid == 0) ctl->fail = 0;
		while (ip < n && !fail) {
			// ---------------- load ----------------
			const uint8_t *g0 = s + ip;
			const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g0) & 15u);
			const uint32_t wlen = n - ip < WIN ? n - ip : WIN;
			const bool last = ip + wlen == n;
			const uint32_t nvec = (mis + wlen + 15u) >> 4;
			dbar();   // the previous window (or block) is no longer read
			for (uint32_t v = tid; v < nvec; v += NT) cp_async16(cwbuf + v * 16, g0 - mis + v * 16);
			cp_async_commit();
			cp_async_wait<0>();
			dbar();
			const uint8_t *cw = cwbuf + mis;
			V3_MARK(P_LOAD);
			V3_COUNT(P_WINDOWS, 1);

			// ---------------- parse: fixed point of "my entry = my left neighbour's exit" ----------------
			const uint32_t seg_lo = static_cast<uint32_t>(tid) * SEG;
			const uint32_t seg_hi = seg_lo + SEG < wlen ? seg_lo + SEG : wlen;
			uint32_t g = seg_lo < wlen ? seg_lo : wlen;
			Walk r = walk_segment<false>(cw, g, seg_hi, wlen, last, 0, 0, nullptr, nullptr, nullptr);
			for (;;) {
				ctl->xs[tid] = r.x;
				dbar();
				const uint32_t ng = tid ? ctl->xs[tid - 1] : 0u;
				const bool changed = ng != g;
				if (changed) {
					g = ng;
					r = walk_segment<false>(cw, g, seg_hi, wlen, last, 0, 0, nullptr, nullptr, nullptr);
				}
				V3_COUNT(P_ITERS, 1);
				if (!dbar_or(changed)) break;
			}
			V3_MARK(P_PARSE);
			// ---------------- block-wide scan of sequence counts and output bytes ----------------
			uint32_t ic = r.c, io = r.o;
#pragma unroll
			for (int sft = 1; sft < 32; sft <<= 1) {
				const uint32_t a = __shfl_up_sync(FULL_MASK, ic, sft), c2 = __shfl_up_sync(FULL_MASK, io, sft);
				if (lane >= sft) { ic += a; io += c2; }
			}
			if (lane == 31) { ctl->wc[warp] = ic; ctl->wo[warp] = io; }
			const bool anybad = dbar_or(r.st == W_BAD);
			for (int w = 0; w < warp; w++) { ic += ctl->wc[w]; io += ctl->wo[w]; }
			// the window takes as many leading segments as fit the token table
			const bool use = ic <= NTOK;
			const uint32_t nuse = dbar_popc(use);
			if (nuse && tid == static_cast<int>(nuse) - 1) {
				ctl->T = ic;
				ctl->O = io;
				ctl->wend = r.x;
				ctl->fin_upto = 0;
				bout[(ic + 31) >> 5] = op + io;   // sentinel: output position at the end of the window
			}
			dbar();
			const uint32_t T = ctl->T, O = ctl->O, wend = ctl->wend;
			V3_MARK(P_SCAN);
			if (anybad || nuse == 0 || wend == 0 || T == 0 || O > cap - op) { fail = true; break; }
			// ---------------- emit ----------------
			if (use && r.c) walk_segment<true>(cw, g, seg_hi, wlen, last, ic - r.c, op + io - r.o, tokpos, bout, out);
			dbar();
			V3_MARK(P_EMIT);
			// ---------------- match ----------------
			const uint32_t nb = (T + 31) >> 5;
			// `last` for the re-parse: the final literal-only sequence is the one whose literals end at wlen
			for (uint32_t k = warp; k < nb; k += NW) match_batch(out, cw, tokpos, bout, k, T, wlen, last, ctl, lane, prof);
			dbar();
			V3_MARK(P_MATCH);
			if (ctl->fail) { fail = true; break; }
			ip += wend;
			op += O;
		}
		if (!fail) {
			// ---------------- flush ----------------
			if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
				const uint32_t nvec = op >> 4;
				const uint4 *o4 = reinterpret_cast<const uint4 *>(out);
				uint4 *g4 = reinterpret_cast<uint4 *>(o);
				for (uint32_t v = tid; v < nvec; v += NT) g4[v] = o4[v];
				const uint32_t tb = nvec << 4;
				if (static_cast<uint32_t>(tid) < (op & 15u)) o[tb + tid] = out[tb + tid];
			} else {
				for (uint32_t i = tid; i < op; i += NT) o[i] = out[i];
			}
			if (tid == 0) {
				status[b].code = LZ4B200_ST_OK;
				status[b].out_len = op;
				status[b].err_pos = 0;
				status[b].aux = 0;
			}
			V3_MARK(P_FLUSH);
		} else {
			// ---------------- exact routine (owns every error message) ----------------
			dbar();
			if (warp == 0) process_block<false>(src, o, d, d.dst_cap, d.hist_avail, status + b, lane);
			V3_COUNT(P_FALLBACK, 1);
			V3_MARK(P_EXACT);
		}
		dbar();   // the output window is free again
		V3_COUNT(P_BLOCKS, 1);
	}
	if (prof) {
		dbar();
		if (tid < PROF_N) atomicAdd(gprof + tid, ctl->prof[tid]);
	}
#undef V3_MARK
#undef V3_COUNT
}

// The hash role: quad q hashes the payload of block first + q.
__device__ __forceinline__ void hash_role(const uint8_t *__restrict__ src, uint32_t first, uint32_t cnt,
					  const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status,
					  uint32_t &computed, uint32_t &declared, bool &want)
{
	const int lane = threadIdx.x & 31, q = lane >> 2;
	want = false;
	computed = declared = 0;
	const uint8_t *s = src;
	uint32_t n = 0;
	bool mine = false;
	if (static_cast<uint32_t>(q) < cnt) {
		const lz4b200_blk_desc d = desc[first + q];
		mine = !(d.flags & LZ4B200_BLK_NOT_K1);
		want = mine && (d.flags & LZ4B200_BLK_HAS_CHECKSUM);
		if (want) {
			s = src + d.src_off;
			n = d.src_len;
		}
	}
	const uint32_t h = quad_xxh32<true, false>(s, n, lane);
	if (want) {
		const uint8_t *t = s + n;
		declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
		computed = h;
	}
	if (mine && (lane & 3) == 0) {
		status[first + q].xxh32_computed = computed;
		status[first + q].xxh32_declared = declared;
	}
}

}  // namespace v3
}  // namespace lz4b200
