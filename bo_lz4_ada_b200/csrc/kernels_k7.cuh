// kernels_k7.cuh -- K7: one CTA per chain (a linked frame, a 1 - 8 MiB block, an exact-placement retry) with NO serial
// walk over the stream -- neither in the parse nor in the copies.
//
// K4 / K6 measured (profiles/r02_k4_*, r02_k6_*): one warp parsing tokens in order and one warp copying matches in
// dependency rounds both end up at ~100 MB/s per stream -- ~170 cycles per sequence, the issue latency of one warp
// times the instructions of a sequence.  A stream is serial twice, and K7 breaks both chains with work instead of
// waiting:
//
//   parse    the compressed bytes are staged in shared memory 8 KiB at a time (a step).  All threads side by side
//            fill a next-token table for the plain tokens (both nibbles below 15: 3 + the literal nibble bytes on).
//            Every thread then owns a 32-byte segment: it guesses its entry point by running up from two segments
//            earlier (a guessed path soon falls in step with the real token chain: on text it meets it inside one
//            segment 56 times out of 100), walks its segment, keeps the positions it visited as a bit mask and
//            where it left the segment.  The true chain is then threaded through the segments: first every segment
//            takes the exit of the one in front of it as its entry point and walks again from there until it meets
//            the path it knows (a few tokens) -- as many iterations as the longest run of segments whose guess and
//            truth do not meet; then the path from segment 0 is marked by pointer jumping over "the segment my exit
//            lands in" and the segments on it are checked against the entry points that path really gives them
//            (behind a long literal run the neighbour is a guess made inside the literals) -- one or two iterations
//            more; 8 in all on text.  Decompress_Sequence's length arithmetic, lib/lz4ada.adb:737-777, is all a walk
//            needs.  tests/test_k7_model_cpu.py is this scheme in Python, checked against the real token chain
//   place    sequence lengths summed per segment, one scan over the CTA: every sequence knows its output position
//   window   the output is built 16 KiB at a time in shared memory.  Every thread writes its own sequences: literal
//            bytes are final at once (Write_Output, :790-824); a match byte whose source lies in front of the window
//            is read from global memory -- final too; a match byte whose source lies inside the window becomes a
//            16-bit POINTER to that byte (Output_With_History, :845-904: byte i of a match is byte i - offset)
//   resolve  pointer jumping over the window: a byte whose pointer target is final takes its value, otherwise it
//            takes the target's pointer.  The depth of the dependency chains (on text one level per ~8 sequences,
//            ~800 deep per 64 KiB) costs log2(depth) rounds of the whole CTA instead of depth rounds of one warp;
//            overlapping matches (offset < length, RLE) are the same thing -- a byte pointing `offset` back
//   flush    the window goes to global memory as whole 16-byte granules, coalesced
//
// Long sequences (>= 64 literal or match bytes) are initialised by a whole warp instead of a thread; sequences whose
// lengths do not fit the staged bytes (runs of hundreds of KiB), stored blocks and everything unusual (offset 0, a
// match reaching before the frame, capacity, truncation) leave the fast path: giant sequences are copied in global
// memory by one warp (the v2 routines), errors hand the rest of the chain to the exact routine (process_block:
// lib/lz4ada.adb:716-904 semantics), as K4 and K6 do.  Block checksums (Check_Checksum, :698-707) are computed by a
// kernel of their own in front of this one (one chain per block, all blocks at once) and only compared here.
#pragma once

#include "kernels_v2.cuh"

namespace lz4b200 {
namespace k7 {

#ifndef LZ4B200_K7_T
#define LZ4B200_K7_T 256
#endif
constexpr uint32_t T = LZ4B200_K7_T;           // threads per CTA (a power of two)
constexpr uint32_t LOG_T = T == 512 ? 9 : T == 256 ? 8 : 7;
constexpr uint32_t WARPS = T / 32;
#ifndef LZ4B200_K7_S
#define LZ4B200_K7_S 32
#endif
constexpr uint32_t S = LZ4B200_K7_S;           // compressed bytes per thread and step (32 or 64: the visited-token mask is one or two registers)
typedef unsigned long long mask_t;
__device__ __forceinline__ uint32_t first_bit(mask_t m) { return static_cast<uint32_t>(__ffsll(static_cast<long long>(m))) - 1u; }
constexpr uint32_t STEP = T * S;               // 8 KiB
constexpr uint32_t SLACK = 1024;               // staged beyond the step: tokens of the last segments complete in here
constexpr uint32_t IN_BYTES = STEP + SLACK + 32;
#ifndef LZ4B200_K7_NWIN
#define LZ4B200_K7_NWIN 16384
#endif
constexpr uint32_t NWIN = LZ4B200_K7_NWIN;     // output window (pointers are 16 bits)
#ifndef LZ4B200_K7_LONG
#define LZ4B200_K7_LONG 64
#endif
constexpr uint32_t LONG = LZ4B200_K7_LONG;                  // literal runs / matches from here on are initialised by a warp
// where a walk left its segment: the block position of the next token, or
constexpr uint32_t EX_END = 0xffffffffu;       // the block's last sequence has been consumed
constexpr uint32_t EX_ERR = 0xfffffffeu;       // a token that cannot be (the exact routine says why)
constexpr uint32_t EX_STOP = 0x80000000u;      // | position: that token does not complete inside the staged bytes -- the next step starts at it

struct Shared {
	mask_t mask[T];
	uint32_t exitp[T], obase[T];
	uint32_t incoming[T], reach[T];
	uint16_t jmp[2][T];
	uint32_t wsum[WARPS], longs[WARPS];
	uint32_t bc[8];
	uint8_t nxt[T * (S + 4)];   // next-token table of the step: segment t, byte o at t * (S + 4) + o (a padded row per segment:
	                            // the lanes of a warp read their segments in different banks)
	__align__(16) uint16_t ptr[NWIN];
	__align__(16) uint8_t win[NWIN];
	__align__(16) uint8_t in[IN_BYTES + 16];   // (+ 16: the literal copy reads up to seven bytes beyond a run)
};

// (statistics, read by lz4b200_decode_linked under LZ4B200_K7_DEBUG: blocks finished, blocks given up, last reason, its block)
__device__ unsigned long long g_prof[16];   // (LZ4B200_K7_DEBUG: steps, parse iterations, resolve calls, resolve rounds, cycles per phase)
__device__ uint32_t g_stats[8];
__device__ __forceinline__ void note_fail(uint32_t why, uint32_t blk, uint32_t at)
{
	if (threadIdx.x == 0) {
		atomicAdd(&g_stats[1], 1u);
		g_stats[2] = why;
		g_stats[3] = blk;
		g_stats[4] = at;
	}
}

struct Tok {
	uint32_t lit, ml, lp, nxt;   // literal bytes at block position lp, match bytes (0: the final literal-only sequence), next token
};
enum : uint32_t { TK_OK = 0, TK_UNSTAGED = 1, TK_ERR = 2 };

// The token at block position x (x < hi) out of the staged bytes: byte at block position p = in[p + bias].  hi = end
// of what is staged (<= n, the block's length).  Decompress_Sequence's length arithmetic, lib/lz4ada.adb:737-777.
__device__ __forceinline__ uint32_t parse_tok(const uint8_t *in, uint32_t bias, uint32_t x, uint32_t n, uint32_t hi, Tok &t)
{
	const uint32_t tk = in[x + bias];
	uint32_t lit = tk >> 4, ml = tk & 15u, p = x + 1;
	if (lit == 15u) {
		for (;;) {
			if (p >= hi) return p >= n ? TK_ERR : TK_UNSTAGED;
			const uint32_t e = in[p + bias];
			p++;
			lit += e;
			if (e != 255u) break;
		}
	}
	t.lit = lit;
	t.lp = p;
	if (lit > n - p) return TK_ERR;
	const uint32_t q = p + lit;
	if (q == n) {   // the final sequence: literals only (:752-764)
		if (ml) return TK_ERR;
		if (q > hi) return TK_UNSTAGED;
		t.ml = 0;
		t.nxt = n;
		return TK_OK;
	}
	if (q + 2 > n) return TK_ERR;
	if (q + 2 > hi) return TK_UNSTAGED;
	p = q + 2;
	if (ml == 15u) {
		for (;;) {
			if (p >= hi) return p >= n ? TK_ERR : TK_UNSTAGED;
			const uint32_t e = in[p + bias];
			p++;
			ml += e;
			if (e != 255u) break;
		}
	}
	t.ml = ml + 4u;
	t.nxt = p;
	return TK_OK;
}

// One chain by one CTA of T threads.  Returns (in every thread) the chain-relative index of the first block the fast
// path did not finish, 0xffffffff when it finished them all.
// pos0: the chain's first block continues a frame whose last pos0 bytes lie in front of `out + pos0` (the single-block
// path under Update: the device-resident history window); 0 for a chain that starts with its frame.
__device__ __forceinline__ uint32_t run_chain(Shared &sh, const lz4b200_chain &ch, const uint8_t *__restrict__ src, uint8_t *out,
					      const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, uint64_t pos0 = 0)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

	// ---- the window (the same values in every thread).  y = frame-relative output position + ph: out_a[y] is the
	// byte and out_a is 16-byte aligned; window element e is y = wy0 + e
	uint8_t *out_a = out;
	uint32_t ph = 0;
	uint32_t wy0 = 0;       // multiple of NWIN
	uint32_t valid_y = 0;   // the window holds the frame's bytes from here on; below: global memory only
	uint32_t res_y = 0;     // final and in global memory below this
	uint32_t init_y = 0;    // initialised below this (res_y <= init_y <= wy0 + NWIN)

	// final values for window elements [lo, hi): pointer jumping.  Ends with every thread past a barrier.
	long long t_mark = 0;   // (profiling: set where a phase begins, read behind the barrier that ends it)
	// Final values for the window elements initialised since the last call ([lo, hi); everything else is final).
	// sh.ptr[i] = 0: element i is final; otherwise the byte is the same as the one ptr[i] elements in front of it
	// (a match byte holds its offset).  First the elements whose source lies in front of the window -- in global memory,
	// final: fetched, all loads of a group in flight together.  Then pointer jumping over the rest: an element takes
	// the byte of its target when that is final, otherwise it adds the target's distance to its own (sums stay below
	// the window size: no source in front of the window is left).  A final byte is written before the pointer that
	// says so (__threadfence_block between them; a reader loads the pointer first).  Ends behind a barrier.
	auto resolve = [&](uint32_t lo, uint32_t hi) {
		__syncthreads();   // the initialisation is complete
		if (tid == 0) {
			atomicAdd(&g_prof[10], static_cast<unsigned long long>(clock64() - t_mark));
			t_mark = clock64();
		}
		if (lo >= hi) return;
		const uint32_t g_lo = lo >> 3, g_hi = (hi + 7u) >> 3;
		const uint32_t s0 = (valid_y > wy0 ? valid_y : wy0) - wy0;   // first element a pointer may name
		for (uint32_t g = g_lo + tid; g < g_hi; g += T) {
			const uint4 P = *reinterpret_cast<const uint4 *>(&sh.ptr[g * 8u]);
			if ((P.x | P.y | P.z | P.w) == 0u) continue;
			const uint32_t w[4] = {P.x, P.y, P.z, P.w};
			uint32_t off[8], v[8];
			bool any = false;
#pragma unroll
			for (int e = 0; e < 8; e++) {
				off[e] = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
				const uint32_t i = g * 8u + e;
				const bool far = off[e] > i - s0 || i < s0;   // (i < s0: not part of this frame's window, pointer 0)
				if (far && off[e]) {
					v[e] = __ldcg(out_a + wy0 + i - off[e]);
					any = true;
				} else {
					v[e] = 0x100u;
				}
			}
			if (!any) continue;
#pragma unroll
			for (int e = 0; e < 8; e++)
				if (v[e] < 0x100u) {
					sh.win[g * 8u + e] = static_cast<uint8_t>(v[e]);
					off[e] = 0;
				}
			*reinterpret_cast<uint4 *>(&sh.ptr[g * 8u]) = make_uint4(off[0] | (off[1] << 16), off[2] | (off[3] << 16), off[4] | (off[5] << 16), off[6] | (off[7] << 16));
		}
		__syncthreads();
		for (;;) {
			uint32_t pending = 0;
			for (uint32_t g = g_lo + tid; g < g_hi; g += T) {
				const uint4 P = *reinterpret_cast<const uint4 *>(&sh.ptr[g * 8u]);
				if ((P.x | P.y | P.z | P.w) == 0u) continue;
				const uint32_t w[4] = {P.x, P.y, P.z, P.w};
				uint32_t off[8], t[8], pj[8], v[8];
#pragma unroll
				for (int e = 0; e < 8; e++) {
					off[e] = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
					t[e] = g * 8u + e - off[e];   // (a final element looks at itself)
				}
#pragma unroll
				for (int e = 0; e < 8; e++) pj[e] = sh.ptr[t[e]];
				__threadfence_block();   // pointers first, bytes second
#pragma unroll
				for (int e = 0; e < 8; e++) v[e] = sh.win[t[e]];
				uint32_t np[8];
#pragma unroll
				for (int e = 0; e < 8; e++) {
					if (off[e] && pj[e] == 0u) sh.win[g * 8u + e] = static_cast<uint8_t>(v[e]);
					np[e] = pj[e] ? off[e] + pj[e] : 0u;
					pending |= np[e];
				}
				__threadfence_block();
				*reinterpret_cast<uint4 *>(&sh.ptr[g * 8u]) = make_uint4(np[0] | (np[1] << 16), np[2] | (np[3] << 16), np[4] | (np[5] << 16), np[6] | (np[7] << 16));
			}
			if (tid == 0) atomicAdd(&g_prof[3], 1ull);
			if (!__syncthreads_or(pending != 0u)) break;
		}
		if (tid == 0) atomicAdd(&g_prof[2], 1ull);
	};
	// window elements [lo, hi) to global memory
	auto flush = [&](uint32_t lo, uint32_t hi) {
		for (uint32_t g = (lo >> 4) + tid; g < ((hi + 15u) >> 4); g += T) {
			const uint32_t a = g << 4;
			uint8_t *gp = out_a + wy0 + a;
			if (a >= lo && a + 16u <= hi) {
				*reinterpret_cast<uint4 *>(gp) = *reinterpret_cast<const uint4 *>(&sh.win[a]);
			} else {
				for (uint32_t k = 0; k < 16u; k++)
					if (a + k >= lo && a + k < hi) gp[k] = sh.win[a + k];
			}
		}
	};
	// everything initialised so far becomes final output
	auto finish_window = [&]() {
		resolve(res_y - wy0, init_y - wy0);
		const long long t1 = clock64();
		flush(res_y - wy0, init_y - wy0);
		res_y = init_y;
		__syncthreads();
		if (tid == 0) {
			atomicAdd(&g_prof[8], static_cast<unsigned long long>(t1 - t_mark));
			atomicAdd(&g_prof[9], static_cast<unsigned long long>(clock64() - t1));
			t_mark = clock64();
		}
	};
	// every element "final" until the initialisation says otherwise (literal bytes and bytes fetched from global memory
	// never touch their pointer).  Between barriers.
	auto fill_ptr = [&]() {
		for (uint32_t g = tid; g < NWIN / 8u; g += T) *reinterpret_cast<uint4 *>(&sh.ptr[g * 8u]) = make_uint4(0u, 0u, 0u, 0u);
	};
	// the output continues at y in global memory (a new frame, or after a copy that went around the window)
	auto restart_window = [&](uint32_t y) {
		wy0 = y & ~(NWIN - 1u);
		valid_y = res_y = init_y = y;
		fill_ptr();
	};

	uint64_t pos = pos0, frame_start = 0;   // chain-relative
	uint32_t ring = 0;                      // Output_Pos of the reference's Buffer (LZ4B200_BLK_RING_CAP blocks)
	uint32_t fail_block = 0xffffffffu;
	ph = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(out) & 15u);
	out_a = out - ph;
	restart_window(static_cast<uint32_t>(pos0) + ph);
	__syncthreads();

	for (uint32_t i = 0; i < ch.n_blocks; i++) {
		const uint32_t b = ch.first_block + i;
		const lz4b200_blk_desc d = desc[b];
		if (d.flags & LZ4B200_BLK_FIRST_OF_FRAME) {
			// (the previous frame is complete: every block ends with finish_window)
			frame_start = pos;
			ring = 0;
			ph = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(out + frame_start) & 15u);
			out_a = out + frame_start - ph;
			restart_window(ph);
		}
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
		bool okay = !(d.flags & LZ4B200_BLK_HASH_ONLY) && fpos0 + cap < 0x7ff00000ull && !(stored && n > cap) && n < 0x7ff00000u && n > 0;
		if (!okay) note_fail(1, b, 0);
		uint32_t computed = 0, declared = 0;
		if (okay && (d.flags & LZ4B200_BLK_HAS_CHECKSUM)) {   // Check_Checksum before any decoding, :672-676 (hashed by the kernel in front)
			computed = status[b].xxh32_computed;
			declared = status[b].xxh32_declared;
			okay = computed == declared;
			if (!okay) note_fail(2, b, 0);
		}
		const uint32_t f0 = static_cast<uint32_t>(fpos0);   // frame-relative position where the block's output starts
		uint32_t fpos = f0;
		if (okay && stored) {
			// stored block (lib/lz4ada.adb:685-695): around the window, one warp
			if (warp == 0) warp_copy<true>(out + pos, s, n, lane);
			fpos += n;
			restart_window(fpos + ph);
			__syncthreads();
		} else if (okay) {
			const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(s) & 15u);
			const uint8_t *sa = s - mis;                    // aligned coordinate = block position + mis
			const uint32_t na16 = (n + mis + 15u) & ~15u;   // end of the block's granules
			uint32_t ip = 0;
			while (ip < n && okay) {
				// ---- stage [ip, ip + STEP + SLACK) ----
				const long long tp0 = clock64();
				const uint32_t base16 = (ip + mis) & ~15u;
				for (uint32_t k = tid * 16u; k < IN_BYTES && base16 + k < na16; k += T * 16u) cp_async16(sh.in + k, sa + base16 + k);
				cp_async_commit();
				cp_async_wait<0>();
				__syncthreads();
				const uint32_t bias = mis - base16;
				const long long tp1 = clock64();
				const uint32_t hi = n < base16 + IN_BYTES - mis ? n : base16 + IN_BYTES - mis;
				const uint32_t x_t = ip + tid * S, seg_end = x_t + S;

				// ---- parse: "where is the next token if one starts here" for every byte of the step, all threads side by
				// side (consecutive lanes take consecutive bytes); 255 = look again when a walk gets there ----
				// (only the plain token -- both nibbles below 15, the sequence inside the staged bytes and not the block's
				// last -- is worked out here: token + literals + offset = 3 + the literal nibble.  Everything else is 255
				// and parsed by the walk that gets there: filling those in here as well cost 4x the cycles, every warp
				// taking the slow path for the one lane in eight that needs it)
				{
					const uint32_t lim = n < hi ? n : hi;
					for (uint32_t k = tid; k < STEP; k += T) {
						const uint32_t x = ip + k;
						uint32_t d = 255u;
						if (x < lim) {
							const uint32_t tk = sh.in[x + bias];
							const uint32_t step = 3u + (tk >> 4);
							if ((tk >> 4) != 15u && (tk & 15u) != 15u && x + step <= lim) d = step;
						}
						sh.nxt[(k / S) * (S + 4u) + (k % S)] = static_cast<uint8_t>(d);
					}
				}
				__syncthreads();
				const long long tq1 = clock64();
				// every thread walks its segment from its first byte; then the true chain is threaded through
				mask_t my_mask = 0;
				uint32_t my_exit = EX_ERR;
				const uint8_t *my_row = sh.nxt + tid * (S + 4u);
				auto walk = [&](uint32_t x0, bool merge) {
					mask_t m = 0;
					uint32_t x = x0, ex;
					for (;;) {
						if (x >= n) { ex = EX_END; break; }
						if (x >= seg_end) { ex = x; break; }
						if (merge && ((my_mask >> (x - x_t)) & 1u) != 0) {   // on the path this thread knows already
							m |= my_mask & ~((mask_t(1) << (x - x_t)) - 1u);
							ex = my_exit;
							break;
						}
						uint32_t d = my_row[x - x_t];
						if (d == 255u) {
							Tok tk;
							const uint32_t c = parse_tok(sh.in, bias, x, n, hi, tk);
							if (c == TK_UNSTAGED) { ex = EX_STOP | x; break; }
							if (c == TK_ERR) { ex = EX_ERR; break; }
							d = tk.nxt - x;
						}
						m |= mask_t(1) << (x - x_t);
						x += d;
					}
					my_mask = m;
					my_exit = ex;
				};
				const bool inside = x_t < n;
				// the guess: not "a token starts at my first byte" (true one time in four) but where a walk that started
				// two segments earlier enters my segment -- by then it has met the real chain 9 times out of 10, so most
				// segments start out right and the iterations below are few.  (The table makes the run-up cheap.)
				uint32_t ent = x_t;
				if (inside && tid > 0) {
					uint32_t p = tid >= 2 ? x_t - 2u * S : ip;
					while (p < x_t) {
						const uint32_t k = p - ip;
						uint32_t d = sh.nxt[(k / S) * (S + 4u) + (k % S)];
						if (d == 255u) {
							Tok tk;
							if (parse_tok(sh.in, bias, p, n, hi, tk) != TK_OK) { p = x_t; break; }   // (no guess: my first byte then)
							d = tk.nxt - p;
						}
						p += d;
					}
					ent = p;
				}
				if (inside) {
					if (ent < seg_end) {
						walk(ent, false);
					} else {
						my_mask = 0;      // the run-up jumps over my segment: nothing to say until a neighbour's exit lands here
						my_exit = ent;
						ent = 0xffffffffu;
					}
				}
				__syncthreads();
				const long long tq2 = clock64();
				// (1) neighbours: a segment's entry point is where the one in front of it leaves off.  No segment is ever
				// dropped here -- a guessed path that jumps far (a literal byte read as a token) must not silence the segments
				// it jumps over, they are most likely right -- so this settles in as many iterations as the longest run of
				// segments whose guessed and real paths do not meet (a dozen on text)
				for (;;) {
					sh.exitp[tid] = my_exit;
					__syncthreads();
					bool changed = false;
					if (tid > 0 && inside) {
						const uint32_t inc = sh.exitp[tid - 1];
						if (inc < EX_STOP && inc >= x_t && inc < seg_end && inc != ent) {
							walk(inc, true);
							ent = inc;
							changed = true;
						}
					}
					if (tid == 0) atomicAdd(&g_prof[1], 1ull);
					if (!__syncthreads_or(changed ? 1 : 0)) break;
				}
				// (2) the chain itself: which segments does the path from thread 0 really pass through (pointer jumping over
				// "the segment my exit lands in"), and does each of them start where that path enters it?  Those that do not
				// (behind a long literal run their neighbour is a guess inside the literals) walk again; the rest keeps to
				// rule (1).  At the fixed point the segments on the path from thread 0 hold the block's true tokens.
				const long long tq3 = clock64();
				bool active = false;
				uint32_t link = T;
				for (;;) {
					link = inside && my_exit < EX_STOP && my_exit - ip < STEP ? (my_exit - ip) / S : T;
					sh.exitp[tid] = my_exit;
					sh.jmp[0][tid] = static_cast<uint16_t>(link);
					sh.reach[tid] = tid == 0 ? 1u : 0u;
					sh.incoming[tid] = EX_ERR;
					__syncthreads();
#pragma unroll 1
					for (uint32_t lvl = 0; lvl < LOG_T; lvl++) {
						const uint32_t j = sh.jmp[lvl & 1u][tid];
						if (j < T && sh.reach[tid]) sh.reach[j] = 1u;
						sh.jmp[(lvl & 1u) ^ 1u][tid] = j < T ? sh.jmp[lvl & 1u][j] : static_cast<uint16_t>(T);
						__syncthreads();
					}
					active = sh.reach[tid] != 0;
					if (active && link < T) sh.incoming[link] = my_exit;
					__syncthreads();
					bool changed = false;
					if (tid > 0 && inside) {
						const uint32_t inc = active ? sh.incoming[tid] : sh.exitp[tid - 1];
						if (inc < EX_STOP && inc >= x_t && inc < seg_end && inc != ent) {
							walk(inc, true);
							ent = inc;
							changed = true;
						}
					}
					if (tid == 0) atomicAdd(&g_prof[1], 1ull);
					if (!__syncthreads_or(changed ? 1 : 0)) break;
				}
				const long long tp2 = clock64();
				// the last segment of the path says where the chain goes on
				if (active && link == T) sh.bc[4] = my_exit;
				if (!active) my_mask = 0;
				else my_mask &= ~((mask_t(1) << (ent - x_t)) - 1u);   // (tokens of the guess in front of the real entry point)

				// ---- place: output bytes per segment, scan ----
				uint32_t tot = 0;
				bool has_long = false;
				for (mask_t bits = my_mask; bits; bits &= bits - 1u) {
					Tok tk;
					parse_tok(sh.in, bias, x_t + first_bit(bits), n, hi, tk);
					tot += tk.lit + tk.ml;
					has_long = has_long || tk.lit >= LONG || tk.ml >= LONG;
				}
				uint32_t incl = tot;
#pragma unroll
				for (int sft = 1; sft < 32; sft <<= 1) {
					const uint32_t v = __shfl_up_sync(FULL_MASK, incl, sft);
					if (lane >= static_cast<uint32_t>(sft)) incl += v;
				}
				const uint32_t lb = __ballot_sync(FULL_MASK, has_long);
				if (lane == 31) sh.wsum[warp] = incl;
				if (lane == 0) sh.longs[warp] = lb;
				if (tid == 0) sh.bc[0] = 0;   // set by whoever meets a match the fast path does not take
				__syncthreads();
				uint32_t before = 0, step_total = 0;
#pragma unroll
				for (uint32_t w = 0; w < WARPS; w++) {
					const uint32_t v = sh.wsum[w];
					if (w < warp) before += v;
					step_total += v;
				}
				const uint32_t step_exit = sh.bc[4];
				if (step_total > cap - (fpos - f0)) { okay = false; note_fail(3, b, ip); break; }   // the exact routine reports the overflow (:678-680)
				const uint32_t my_o = fpos + before + incl - tot;
				sh.obase[tid] = my_o;
				sh.mask[tid] = my_mask;
				__syncthreads();
				const long long tp3 = clock64();
				t_mark = tp3;
				if (tid == 0) {
					atomicAdd(&g_prof[0], 1ull);
					atomicAdd(&g_prof[4], static_cast<unsigned long long>(tp1 - tp0));
					atomicAdd(&g_prof[11], static_cast<unsigned long long>(tq1 - tp1));
					atomicAdd(&g_prof[12], static_cast<unsigned long long>(tq2 - tq1));
					atomicAdd(&g_prof[13], static_cast<unsigned long long>(tq3 - tq2));
					atomicAdd(&g_prof[14], static_cast<unsigned long long>(tp2 - tq3));
					atomicAdd(&g_prof[5], static_cast<unsigned long long>(tp2 - tp1));
					atomicAdd(&g_prof[6], static_cast<unsigned long long>(tp3 - tp2));
				}

				// ---- window passes: initialise, and whenever the window is full resolve + flush + move on ----
				uint32_t done_y = fpos + ph;
				const uint32_t end_y = done_y + step_total;
				while (done_y < end_y) {
					const uint32_t hi_y = end_y < wy0 + NWIN ? end_y : wy0 + NWIN;
					const uint32_t lo_f = done_y - ph, hi_f = hi_y - ph;         // frame-relative bounds of this pass
					const uint32_t src_y = valid_y > wy0 ? valid_y : wy0;       // sources from here on are window elements
					// (a) every thread: the short sequences of its segment, up to four bytes per turn of one flat loop (a
					// literal byte is final at once, a match byte is its offset; what lies in front of the window is
					// fetched by resolve)
					{
						mask_t bits = my_mask;
						uint32_t o = my_o, lit_left = 0, m_left = 0, lp = 0, off = 0;
						const uint32_t shift = ph - wy0, span = hi_f - lo_f;   // element of frame position f = f + shift
						for (;;) {
							if ((lit_left | m_left) == 0u) {
								if (!bits || o >= hi_f) break;
								Tok tk;
								parse_tok(sh.in, bias, x_t + first_bit(bits), n, hi, tk);
								bits &= bits - 1u;
								const uint32_t m = o + tk.lit, e = m + tk.ml;
								if (tk.ml) {
									const uint32_t q = tk.lp + tk.lit + bias;
									off = sh.in[q] | (static_cast<uint32_t>(sh.in[q + 1u]) << 8);
									if (off == 0 || off > m) { sh.bc[0] = 1; break; }   // lib/lz4ada.adb:766-772, :864-874
								}
								if (e <= lo_f || tk.lit >= LONG || tk.ml >= LONG) { o = e; continue; }
								lit_left = tk.lit;
								m_left = tk.ml;
								lp = tk.lp + bias;
								if (lit_left == 0u && m_left == 0u) continue;
							}
							// up to four bytes of the run in progress (eight per turn, literals and match bytes under one
							// roof, measured slower: 3.2 against 1.7 G cycles for the initialisation of 512 MiB)
							if (lit_left) {
								const uint32_t k = lit_left < 4u ? lit_left : 4u;
								uint32_t v[4];
#pragma unroll
								for (uint32_t j = 0; j < 4u; j++) v[j] = sh.in[lp + j];   // (a few bytes beyond a run: staged)
#pragma unroll
								for (uint32_t j = 0; j < 4u; j++)
									if (j < k && o + j - lo_f < span) sh.win[o + j + shift] = static_cast<uint8_t>(v[j]);
								lp += k;
								o += k;
								lit_left -= k;
							} else {
								const uint32_t k = m_left < 4u ? m_left : 4u;
#pragma unroll
								for (uint32_t j = 0; j < 4u; j++)
									if (j < k && o + j - lo_f < span) sh.ptr[o + j + shift] = static_cast<uint16_t>(off);
								o += k;
								m_left -= k;
							}
						}
					}
					// (b) every warp: the long sequences of its 32 segments, a lane per byte
					for (uint32_t lb2 = sh.longs[warp]; lb2; lb2 &= lb2 - 1u) {
						const uint32_t tt = warp * 32u + __ffs(lb2) - 1u;
						uint32_t o = sh.obase[tt];
						const uint32_t xs = ip + tt * S;
						for (mask_t bits = sh.mask[tt]; bits && o < hi_f; bits &= bits - 1u) {
							Tok tk;
							parse_tok(sh.in, bias, xs + first_bit(bits), n, hi, tk);
							const uint32_t m = o + tk.lit, e = m + tk.ml;
							if (e > lo_f && (tk.lit >= LONG || tk.ml >= LONG)) {
								const uint32_t l0 = o > lo_f ? o : lo_f, l1 = m < hi_f ? m : hi_f;
								for (uint32_t f = l0 + lane; f < l1; f += 32u) sh.win[f + ph - wy0] = sh.in[tk.lp + (f - o) + bias];
								if (tk.ml) {
									const uint32_t q = tk.lp + tk.lit + bias;
									const uint32_t off = sh.in[q] | (static_cast<uint32_t>(sh.in[q + 1u]) << 8);
									if (off == 0 || off > m) { sh.bc[0] = 1; break; }
									const uint32_t m0 = m > lo_f ? m : lo_f, m1 = e < hi_f ? e : hi_f;
									for (uint32_t f = m0 + lane; f < m1; f += 32u) {
										// byte k of a match repeats its first `off` bytes: point straight at the first period
										const uint32_t k = f - m;
										const uint32_t sf = m - off + (k < off ? k : k % off);
										const uint32_t el = f + ph - wy0;
										if (sf + ph >= src_y) sh.ptr[el] = static_cast<uint16_t>(f - sf);
										else sh.win[el] = __ldcg(out_a + sf + ph);
									}
								}
							}
							o = e;
						}
					}
					done_y = init_y = hi_y;
					if (hi_y == wy0 + NWIN) {
						finish_window();
						wy0 += NWIN;
						fill_ptr();
						__syncthreads();
					} else {
						__syncthreads();
						if (tid == 0) {
							atomicAdd(&g_prof[10], static_cast<unsigned long long>(clock64() - t_mark));
							t_mark = clock64();
						}
					}
					if (sh.bc[0]) break;
				}
				if (sh.bc[0]) { okay = false; note_fail(4, b, ip); break; }
				if (tid == 0) atomicAdd(&g_prof[7], static_cast<unsigned long long>(clock64() - tp3));   // (window passes, resolve and flush included)
				fpos += step_total;

				// ---- where the chain goes on ----
				if (step_exit == EX_END) {
					ip = n;
				} else if (step_exit == EX_ERR) {
					okay = false;
					note_fail(5, b, ip);
				} else if (step_exit & EX_STOP) {
					const uint32_t xs = step_exit & ~EX_STOP;
					if (xs != ip) {
						ip = xs;
					} else {
						// a sequence too long for the staged bytes: one warp, in global memory (the v2 routines)
						finish_window();
						if (warp == 0) {
							uint32_t lp = 0, lit = 0, ml = 0, nxt = 0, off = 0;
							bool fine = parse_token_wide(s, n, ip, lp, lit, ml, nxt, lane);
							if (fine && ml) off = ld_u8<true>(s + lp + lit) | (ld_u8<true>(s + lp + lit + 1) << 8);
							const uint32_t left = cap - (fpos - f0);
							if (fine && (lit > left || ml > left - lit || (ml && (off == 0 || off > fpos + lit)))) fine = false;
							if (fine) {
								uint8_t *o = out + frame_start + fpos;
								warp_copy<true>(o, s + lp, lit, lane);
								if (ml) match_copy(o + lit, off, ml, lane);
							}
							if (lane == 0) {
								sh.bc[1] = fine ? 1u : 0u;
								sh.bc[2] = lit + ml;
								sh.bc[3] = nxt;
							}
						}
						__syncthreads();
						if (!sh.bc[1]) {
							okay = false;
							note_fail(6, b, ip);
						} else {
							fpos += sh.bc[2];
							ip = sh.bc[3];
							restart_window(fpos + ph);
						}
						__syncthreads();
					}
				} else {
					ip = step_exit;
				}
			}
			if (okay) finish_window();
		}
		if (!okay) {
			// anything out of the ordinary: the exact routine takes over from this block
			fail_block = i;
			break;
		}
		if (tid == 0) {
			atomicAdd(&g_stats[0], 1u);
			status[b].code = LZ4B200_ST_OK;
			status[b].out_len = fpos - f0;
			status[b].err_pos = 0;
			status[b].aux = 0;
			status[b].xxh32_computed = computed;
			status[b].xxh32_declared = declared;
		}
		pos += fpos - f0;
		ring += fpos - f0;
	}
	__syncthreads();
	return fail_block;
}

}  // namespace k7
}  // namespace lz4b200
