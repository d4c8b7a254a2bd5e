// kernels.cuh -- sm_100a device code of the LZ4 block decoder (K1/K4), XXH32 chains (K3),
// size pre-pass (K5).  No tensor cores: the path is byte/integer work bounded by HBM and by
// instruction issue.
//
// Reference code replaced (behaviour, not structure): lib/lz4ada.adb:661-904 (block decode),
// :923-1026 (XXHash32).  The reference decodes serially into a 64 KiB ring; here every block
// writes a flat output range, so the three-phase history copy of :845-904 collapses into one
// backwards reference into already-written output.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lz4b200.h"

namespace lz4b200 {

constexpr uint32_t FULL_MASK = 0xffffffffu;
// lib/lz4ada.ads:324-328
constexpr uint32_t PRIME_1 = 2654435761u;
constexpr uint32_t PRIME_2 = 2246822519u;
constexpr uint32_t PRIME_3 = 3266489917u;
constexpr uint32_t PRIME_4 = 668265263u;
constexpr uint32_t PRIME_5 = 374761393u;

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int s) { return __funnelshift_l(v, v, s); }

// Rot_Mul, lib/lz4ada.adb:982-985
__device__ __forceinline__ uint32_t xxh_round(uint32_t acc, uint32_t x)
{
	return rotl32(acc + x * PRIME_2, 13) * PRIME_1;
}

// A run of rounds of ONE accumulator is a serial chain, and written as above every round is three dependent
// operations (IMAD, SHF, IMAD: 14 cycles on sm_100a, the multiplier and the shifter sit on different pipes).
// Carrying s = acc + x * P2 instead, a round is s' = rotl(s, 13) * P1 + x' * P2 with the product x' * P2 off the
// chain: two dependent operations (SHF, IMAD).  LZ4B200_XXH_CHAIN = 2 goes one further: rotl(s, 13) = lo(s * 2^13) +
// hi(s * 2^13), so s' = s * (P1 << 13) + hi(s * 2^13) * P1 + x' * P2 -- IMAD.HI, then IMAD, both on the multiplier
// pipe (the 2^13 comes from constant memory so that it stays a multiplication).  Same integers either way.
// Inline PTX because the compiler re-associates the sum and puts x' * P2 back on the chain otherwise.
// Measured per round by tools/probes/xxh_chain_probe.cu.
#ifndef LZ4B200_XXH_CHAIN
#define LZ4B200_XXH_CHAIN 1
#endif
__constant__ uint32_t xxh_two13 = 8192u;

__device__ __forceinline__ uint32_t xxh_step(uint32_t s, uint32_t x)
{
	uint32_t c, r;
	asm("mul.lo.u32 %0, %1, %2;" : "=r"(c) : "r"(x), "r"(PRIME_2));
#if LZ4B200_XXH_CHAIN == 2
	uint32_t hi, a;
	asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(s), "r"(xxh_two13));
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(s), "r"(PRIME_1 << 13), "r"(c));
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(PRIME_1), "r"(a));
#else
	const uint32_t t = rotl32(s, 13);
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(t), "r"(PRIME_1), "r"(c));
#endif
	return r;
}

// acc after N more rounds (Rot_Mul N times, lib/lz4ada.adb:982-985)
template <int N> __device__ __forceinline__ uint32_t xxh_fold(uint32_t acc, const uint32_t (&x)[N])
{
	uint32_t s = acc + x[0] * PRIME_2;
#pragma unroll
	for (int j = 1; j < N; j++) s = xxh_step(s, x[j]);
	return rotl32(s, 13) * PRIME_1;
}

// Loads.  RO = the bytes are never written during this kernel (compressed input, or output of an
// earlier launch): non-coherent path.  Otherwise a plain load, ordered by __syncwarp().
template <bool RO> __device__ __forceinline__ uint32_t ld_u8(const uint8_t *p)
{
	if (RO) return __ldg(p);
	return *p;
}
template <bool RO> __device__ __forceinline__ uint32_t ld_u32(const uint32_t *p)
{
	if (RO) return __ldg(p);
	return *p;
}
template <bool RO> __device__ __forceinline__ uint4 ld_u128(const uint4 *p)
{
	if (RO) return __ldg(p);
	return *p;
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
	asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---------------------------------------------------------------------------------------------
// XXH32 by a quad (4 consecutive lanes): lane `sub` owns accumulator `sub` and consumes word
// `sub` of every 16-byte stripe (XXHash32.Process, lib/lz4ada.adb:979-991).  The chain is serial
// per quad, so a warp carries eight independent chains.  All 32 lanes must call this together
// (quads without work pass n = 0).  Every lane of the quad returns the digest of [p, p + n).
// ---------------------------------------------------------------------------------------------
struct XxhState {            // streaming state, same fields as lib/lz4ada.ads:335-343
	uint32_t acc[4];
	uint8_t  buf[16];
	uint32_t buf_size;
	uint32_t pad;
	uint64_t total_len;
};

__device__ __forceinline__ uint32_t xxh_init_acc(int sub)   // Reset, :932-940 with Seed = 0
{
	return sub == 0 ? PRIME_1 + PRIME_2 : sub == 1 ? PRIME_2 : sub == 2 ? 0u : 0u - PRIME_1;
}

template <bool RO, bool PREFETCH>
__device__ __forceinline__ uint32_t quad_stripes(const uint8_t *p, uint64_t nstripes, uint32_t acc,
						 int sub)
{
	// The chain acc -> acc is ~13 dependent cycles per stripe; the loads are not on it.  Keep the
	// next 8 stripes in flight while the current 8 are folded in (software pipeline), and pull lines
	// into L2 a few KiB ahead so that the in-flight loads see L2 latency, not DRAM latency.
	const uintptr_t a = reinterpret_cast<uintptr_t>(p);
	const uint32_t mis = static_cast<uint32_t>(a & 3);
	const uint32_t *w = reinterpret_cast<const uint32_t *>(a - mis) + sub;
	constexpr uint64_t AHEAD = 8192 / 16;   // stripes of L2 prefetch distance
	const uint32_t sh = mis * 8;
	uint64_t s = 0;
	if (nstripes >= 8) {
		uint32_t cur[8], nxt[8];
		if (mis == 0) {
#pragma unroll
			for (int j = 0; j < 8; j++) cur[j] = ld_u32<RO>(w + j * 4);
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++)
				cur[j] = __funnelshift_r(ld_u32<RO>(w + j * 4), ld_u32<RO>(w + j * 4 + 1), sh);
		}
		for (s = 8; s + 8 <= nstripes; s += 8) {
			if (PREFETCH && sub == 0 && s + AHEAD < nstripes) prefetch_l2(w + (s + AHEAD) * 4);
			if (mis == 0) {
#pragma unroll
				for (int j = 0; j < 8; j++) nxt[j] = ld_u32<RO>(w + (s + j) * 4);
			} else {
#pragma unroll
				for (int j = 0; j < 8; j++)
					nxt[j] = __funnelshift_r(ld_u32<RO>(w + (s + j) * 4), ld_u32<RO>(w + (s + j) * 4 + 1), sh);
			}
			acc = xxh_fold<8>(acc, cur);
#pragma unroll
			for (int j = 0; j < 8; j++) cur[j] = nxt[j];
		}
		acc = xxh_fold<8>(acc, cur);
	}
	for (; s < nstripes; s++) {
		const uint32_t x = mis == 0 ? ld_u32<RO>(w + s * 4)
					    : __funnelshift_r(ld_u32<RO>(w + s * 4), ld_u32<RO>(w + s * 4 + 1), sh);
		acc = xxh_round(acc, x);
	}
	return acc;
}

// Final, lib/lz4ada.adb:993-1017: `tail` points at the < 16 bytes left after the stripes.
template <bool RO>
__device__ __forceinline__ uint32_t xxh_finish(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
					       uint64_t total_len, const uint8_t *tail, uint32_t rem)
{
	uint32_t h = total_len >= 16 ? rotl32(a0, 1) + rotl32(a1, 7) + rotl32(a2, 12) + rotl32(a3, 18)
				     : a2 + PRIME_5;
	h += static_cast<uint32_t>(total_len);
	while (rem >= 4) {
		uint32_t x = ld_u8<RO>(tail) | (ld_u8<RO>(tail + 1) << 8) | (ld_u8<RO>(tail + 2) << 16) |
			     (ld_u8<RO>(tail + 3) << 24);
		h = rotl32(h + x * PRIME_3, 17) * PRIME_4;
		tail += 4;
		rem -= 4;
	}
	while (rem) {
		h = rotl32(h + ld_u8<RO>(tail) * PRIME_5, 11) * PRIME_1;
		tail += 1;
		rem -= 1;
	}
	h = (h ^ (h >> 15)) * PRIME_2;
	h = (h ^ (h >> 13)) * PRIME_3;
	return h ^ (h >> 16);
}

template <bool RO, bool PREFETCH>
__device__ __forceinline__ uint32_t quad_xxh32(const uint8_t *p, uint64_t n, int lane)
{
	const int sub = lane & 3;
	const uint64_t nstripes = n >> 4;
	uint32_t acc = quad_stripes<RO, PREFETCH>(p, nstripes, xxh_init_acc(sub), sub);
	const uint32_t a0 = __shfl_sync(FULL_MASK, acc, 0, 4);
	const uint32_t a1 = __shfl_sync(FULL_MASK, acc, 1, 4);
	const uint32_t a2 = __shfl_sync(FULL_MASK, acc, 2, 4);
	const uint32_t a3 = __shfl_sync(FULL_MASK, acc, 3, 4);
	return xxh_finish<RO>(a0, a1, a2, a3, n, p + (nstripes << 4), static_cast<uint32_t>(n & 15));
}

// ---------------------------------------------------------------------------------------------
// Long spans (content checksums): the chain needs 16 bytes every ~14 cycles per frame, but a frame
// only has four lanes, so register-staged loads keep far too few bytes in flight (measured 1.3 TB/s
// with 4096 chains).  Here each quad streams its span through a 2 KiB shared-memory ring filled by
// cp.async (LDGSTS, 16 bytes per lane, eight 256-byte groups in flight per quad = 8 MB in flight
// for 4096 frames) and folds stripes straight out of shared memory.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t XXH_GROUPS = 8;                   // cp.async groups in flight per quad
// behind each ring: 32 bytes that mirror its first 32 (a batch of stripes read from the last slot runs on linearly
// into them), and 16 bytes of skew so that the eight quads of a warp start in different banks: conflict-free LDS
constexpr uint32_t XXH_RING_PAD = 48;
constexpr uint32_t XXH_GROUP_BYTES = 256;            // 4 x (4 lanes x 16 B): the default group, 2 KiB ring per quad
constexpr uint32_t XXH_RING_BYTES = XXH_GROUPS * XXH_GROUP_BYTES;
constexpr uint32_t XXH_RING_STRIDE = XXH_RING_BYTES + XXH_RING_PAD;
// few, long spans (a handful of 4 MiB frames): the same ring with 1 KiB groups = 8 KiB in flight per quad
constexpr uint32_t XXH_BIG_GROUP_BYTES = 1024;
constexpr uint32_t XXH_BIG_RING_STRIDE = XXH_GROUPS * XXH_BIG_GROUP_BYTES + XXH_RING_PAD;
// a few warps per SM: 512-byte groups = 4 KiB in flight per quad (33 KB per warp, one warp per CTA)
constexpr uint32_t XXH_MID_GROUP_BYTES = 512;
constexpr uint32_t XXH_MID_RING_STRIDE = XXH_GROUPS * XXH_MID_GROUP_BYTES + XXH_RING_PAD;
// ... and when there is at most one warp per SM, 2 KiB groups = 16 KiB in flight per quad (131 KB per warp)
constexpr uint32_t XXH_HUGE_GROUP_BYTES = 2048;
constexpr uint32_t XXH_HUGE_RING_STRIDE = XXH_GROUPS * XXH_HUGE_GROUP_BYTES + XXH_RING_PAD;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// All 32 lanes call this together; ring = this warp's 8 x (8 * GROUP_BYTES + XXH_RING_PAD) bytes of shared memory.
// Reads whole 16-byte granules: at most 19 bytes past p + n (the batch API's 32 bytes of slack cover it).
// The stripes of a span through the ring: acc in, acc out (the chain itself; a caller with a running state -- the
// content checksum under Update -- starts and finishes it elsewhere).
template <uint32_t GROUP_BYTES>
__device__ __forceinline__ uint32_t quad_ring_stripes_t(const uint8_t *p, uint64_t nstripes, uint32_t acc, uint8_t *ring, int lane)
{
	constexpr uint32_t RING_BYTES = XXH_GROUPS * GROUP_BYTES;
	constexpr uint32_t RING_STRIDE = RING_BYTES + XXH_RING_PAD;
	constexpr uint32_t BATCH = GROUP_BYTES / 16;   // stripes per group
	const int sub = lane & 3, q = lane >> 2;
	uint8_t *my_ring = ring + q * RING_STRIDE;
	const uintptr_t a = reinterpret_cast<uintptr_t>(p);
	const uint32_t mis = static_cast<uint32_t>(a & 15);
	const uint8_t *abase = p - mis;                                   // 16-byte aligned
	// bytes from abase the stripes touch, in whole granules (+ 4: the word a misaligned last stripe spills into)
	const uint64_t need = nstripes ? ((mis + (nstripes << 4) + 4 + 15) & ~uint64_t(15)) : 0;
	const uint32_t my_groups = static_cast<uint32_t>((need + GROUP_BYTES - 1) / GROUP_BYTES);
	const uint32_t max_groups = __reduce_max_sync(FULL_MASK, my_groups);
	const uint32_t sh = (mis & 3) * 8;

	auto issue = [&](uint32_t g) {
		if (g < my_groups) {
			const uint64_t gb = static_cast<uint64_t>(g) * GROUP_BYTES;
			const uint32_t so = static_cast<uint32_t>(gb) & (RING_BYTES - 1);
			uint8_t *slot = my_ring + so + sub * 16;
			const uint8_t *gsrc = abase + gb + sub * 16;
			if (gb + GROUP_BYTES <= need) {
#pragma unroll
				for (uint32_t c = 0; c < GROUP_BYTES / 64; c++) cp_async16(slot + c * 64, gsrc + c * 64);
			} else {
				// the span's last group: only the granules that belong to it
#pragma unroll
				for (uint32_t c = 0; c < GROUP_BYTES / 64; c++)
					if (gb + c * 64 + sub * 16 < need) cp_async16(slot + c * 64, gsrc + c * 64);
			}
			// the mirror of the ring's first 32 bytes, behind its end
			if (so == 0 && sub < 2 && gb + sub * 16 < need) cp_async16(my_ring + RING_BYTES + sub * 16, gsrc);
		}
		cp_async_commit();
	};
	for (uint32_t g = 0; g < XXH_GROUPS; g++) issue(g);
	// Iteration g runs once group g has landed and folds stripe batch g - 1 (BATCH stripes = one group):
	// lagging by one group means the bytes a misaligned span spills into the next group are there.
	const uint64_t n_batches = (nstripes + BATCH - 1) / BATCH;
	for (uint32_t g = 0; g <= max_groups; g++) {
		// commits so far: 8 (prologue) + (g - 1); all but the newest 6 are complete => groups 0..g landed
		cp_async_wait<XXH_GROUPS - 2>();
		__syncwarp();
		if (g >= 1 && g - 1 < n_batches) {
			const uint64_t s0 = static_cast<uint64_t>(g - 1) * BATCH;
			// this lane's word of stripe s0; stripe j is 4 j words further on -- linearly, also out of the last slot
			// (whose spill is the mirror), so every load below is base + immediate
			const uint32_t b0 = (((g - 1) * GROUP_BYTES) & (RING_BYTES - 1)) + (mis & ~3u) + (static_cast<uint32_t>(sub) << 2);
			const uint32_t *w = reinterpret_cast<const uint32_t *>(my_ring + b0);
			const uint64_t left = nstripes - s0;
			if (left >= BATCH) {
				// a whole group, unrolled: the loads of the next 16 stripes are scheduled under the chain of these 16
				if (sh == 0) {
#pragma unroll
					for (uint32_t c = 0; c < BATCH / 16; c++) {
						uint32_t x[16];
#pragma unroll
						for (int j = 0; j < 16; j++) x[j] = w[(c * 16 + j) * 4];
						acc = xxh_fold<16>(acc, x);
					}
				} else {
#pragma unroll
					for (uint32_t c = 0; c < BATCH / 16; c++) {
						uint32_t x[16];
#pragma unroll
						for (int j = 0; j < 16; j++) x[j] = __funnelshift_r(w[(c * 16 + j) * 4], w[(c * 16 + j) * 4 + 1], sh);
						acc = xxh_fold<16>(acc, x);
					}
				}
			} else {
				const uint32_t cnt = static_cast<uint32_t>(left);
				uint32_t j0 = 0;
				for (; j0 + 16 <= cnt; j0 += 16) {
					uint32_t x[16];
#pragma unroll
					for (int j = 0; j < 16; j++) {
						x[j] = w[(j0 + j) * 4];
						if (sh) x[j] = __funnelshift_r(x[j], w[(j0 + j) * 4 + 1], sh);
					}
					acc = xxh_fold<16>(acc, x);
				}
				for (uint32_t j = j0; j < cnt; j++) {
					uint32_t x = w[j * 4];
					if (sh) x = __funnelshift_r(x, w[j * 4 + 1], sh);
					acc = xxh_round(acc, x);
				}
			}
		}
		__syncwarp();
		if (g >= 1) issue(g - 1 + XXH_GROUPS);   // refill the slot batch g - 1 has just been read from
	}
	cp_async_wait<0>();
	return acc;
}

template <uint32_t GROUP_BYTES>
__device__ __forceinline__ uint32_t quad_xxh32_stream_t(const uint8_t *p, uint64_t n, uint8_t *ring, int lane)
{
	const uint64_t nstripes = n >> 4;
	const uint32_t acc = quad_ring_stripes_t<GROUP_BYTES>(p, nstripes, xxh_init_acc(lane & 3), ring, lane);
	const uint32_t a0 = __shfl_sync(FULL_MASK, acc, 0, 4);
	const uint32_t a1 = __shfl_sync(FULL_MASK, acc, 1, 4);
	const uint32_t a2 = __shfl_sync(FULL_MASK, acc, 2, 4);
	const uint32_t a3 = __shfl_sync(FULL_MASK, acc, 3, 4);
	return xxh_finish<true>(a0, a1, a2, a3, n, p + (nstripes << 4), static_cast<uint32_t>(n & 15));
}

__device__ __forceinline__ uint32_t quad_xxh32_stream(const uint8_t *p, uint64_t n, uint8_t *ring, int lane)
{
	return quad_xxh32_stream_t<XXH_GROUP_BYTES>(p, n, ring, lane);
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative copies (Write_Output, lib/lz4ada.adb:790-824 -- but exact: nothing is ever
// written outside [dst, dst + n)).  Requires dst - src >= n or disjoint buffers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 shift_combine(const uint4 &A, const uint4 &B, uint32_t ws, uint32_t bs)
{
	uint32_t w0, w1, w2, w3, w4;
	switch (ws) {   // warp-uniform
	case 0: w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; break;
	case 1: w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; break;
	case 2: w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; break;
	default: w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; break;
	}
	return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs),
			  __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
}

template <bool RO>
__device__ __noinline__ void warp_copy_long(uint8_t *dst, const uint8_t *src, uint32_t n, int lane)
{
	// head: bring dst to a 16-byte boundary
	const uint32_t head = (16u - static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
	if (lane < head) dst[lane] = static_cast<uint8_t>(ld_u8<RO>(src + lane));
	dst += head;
	src += head;
	n -= head;
	const uint32_t nvec = n >> 4;
	const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src) & 15);
	const uint4 *s4 = reinterpret_cast<const uint4 *>(src - m);
	uint4 *d4 = reinterpret_cast<uint4 *>(dst);
	uint32_t v = lane;
	if (m == 0) {
		for (; v + 96 < nvec; v += 128) {   // 4 x 512 B in flight per warp
			uint4 r0 = ld_u128<RO>(s4 + v), r1 = ld_u128<RO>(s4 + v + 32);
			uint4 r2 = ld_u128<RO>(s4 + v + 64), r3 = ld_u128<RO>(s4 + v + 96);
			d4[v] = r0; d4[v + 32] = r1; d4[v + 64] = r2; d4[v + 96] = r3;
		}
		for (; v < nvec; v += 32) d4[v] = ld_u128<RO>(s4 + v);
	} else {
		const uint32_t ws = m >> 2, bs = (m & 3) * 8;
		for (; v + 32 < nvec; v += 64) {
			uint4 a0 = ld_u128<RO>(s4 + v), b0 = ld_u128<RO>(s4 + v + 1);
			uint4 a1 = ld_u128<RO>(s4 + v + 32), b1 = ld_u128<RO>(s4 + v + 33);
			d4[v] = shift_combine(a0, b0, ws, bs);
			d4[v + 32] = shift_combine(a1, b1, ws, bs);
		}
		for (; v < nvec; v += 32) d4[v] = shift_combine(ld_u128<RO>(s4 + v), ld_u128<RO>(s4 + v + 1), ws, bs);
	}
	const uint32_t done = nvec << 4, tail = n & 15;
	if (lane < tail) dst[done + lane] = static_cast<uint8_t>(ld_u8<RO>(src + done + lane));
}

template <bool RO>
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int lane)
{
	if (n <= 32) {
		if (lane < n) dst[lane] = static_cast<uint8_t>(ld_u8<RO>(src + lane));
	} else if (n < 128) {
		for (uint32_t i = lane; i < n; i += 32) dst[i] = static_cast<uint8_t>(ld_u8<RO>(src + i));
	} else {
		warp_copy_long<RO>(dst, src, n, lane);
	}
}

// Match copy: dst[0 .. ml) = dst[-offset ...] with LZ4 overlap semantics
// (Output_With_History, lib/lz4ada.adb:845-904, phases I and R; phase H is the same
// backwards read because the output is flat).
__device__ __noinline__ void match_copy_overlap(uint8_t *d, uint32_t offset, uint32_t ml, int lane)
{
	uint32_t done = 0;
	if (offset < 32) {
		// first bytes straight from the pattern: d[i] = pattern[i mod offset]
		const uint32_t m0 = ml < 32 ? ml : 32;
		if (lane < m0) d[lane] = *(d - offset + (static_cast<uint32_t>(lane) % offset));
		done = m0;
		if (done == ml) return;
		__syncwarp();
		if ((offset & (offset - 1)) == 0 && offset <= 16 && ml - done >= 64) {
			// period divides 16: every 16-byte aligned vector of the run is the same
			// register pattern -> store-only replication (RLE runs, zero pages)
			uint8_t *A = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(d) + 15) & ~uintptr_t(15));
			const uint4 V = *reinterpret_cast<const uint4 *>(A);   // A + 16 <= d + 32: materialised
			uint4 *body = reinterpret_cast<uint4 *>(A + 16);
			uint8_t *end = d + ml;
			const uint32_t nvec = static_cast<uint32_t>(end - (A + 16)) >> 4;
			uint32_t v = lane;
			for (; v + 96 < nvec; v += 128) {
				body[v] = V; body[v + 32] = V; body[v + 64] = V; body[v + 96] = V;
			}
			for (; v < nvec; v += 32) body[v] = V;
			uint8_t *T = A + 16 + (static_cast<size_t>(nvec) << 4);
			const uint32_t tail = static_cast<uint32_t>(end - T);
			if (lane < tail) T[lane] = A[lane];
			return;
		}
	}
	// doubling: L = largest multiple of the period not exceeding the bytes already valid
	while (done < ml) {
		const uint32_t avail = offset + done;
		const uint32_t L = avail - (avail % offset);
		const uint32_t chunk = (ml - done) < L ? (ml - done) : L;
		warp_copy<false>(d + done, d + done - L, chunk, lane);
		done += chunk;
		if (done < ml) __syncwarp();
	}
}

__device__ __forceinline__ void match_copy(uint8_t *d, uint32_t offset, uint32_t ml, int lane)
{
	__syncwarp();   // earlier stores of other lanes must be visible to the loads below
	if (offset >= ml) {
		warp_copy<false>(d, d - offset, ml, lane);
	} else {
		match_copy_overlap(d, offset, ml, lane);
	}
}

// Process_Variable_Length, lib/lz4ada.adb:724-735: add bytes until one /= 255.  32 bytes per step
// with a ballot for the first non-255 (z9m has 16 448 0xff bytes in one length).
// Returns false when the extension runs past the end of the block.
__device__ __forceinline__ bool read_length_ext(const uint8_t *src, uint32_t &ip, uint32_t iend,
						uint32_t &len, int lane)
{
	for (;;) {
		const uint32_t idx = ip + lane;
		const uint32_t b = idx < iend ? ld_u8<true>(src + idx) : 0x100u;
		const uint32_t stop = __ballot_sync(FULL_MASK, b != 255u);
		if (stop == 0) {
			len += 255u * 32u;
			ip += 32;
			continue;
		}
		const int first = __ffs(stop) - 1;
		const uint32_t bv = __shfl_sync(FULL_MASK, b, first);
		if (bv == 0x100u) return false;
		len += 255u * first + bv;
		ip += first + 1;
		return true;
	}
}

struct BlockResult {
	uint32_t code;
	uint32_t out_len;
	uint32_t err_pos;
	int32_t  aux;
};

// Decompress_Full_Block / Decompress_Sequence, lib/lz4ada.adb:716-788, one warp per block.
// All lanes walk the token stream in lock-step (warp-uniform control flow, broadcast loads);
// literal and match bytes are moved by the whole warp.
// `o` = where this block's output starts, `hist` = bytes before `o` that belong to the same frame.
// ALLOW_HIST = the bytes before `o` are already final (linked frames decoded in order).
template <bool ALLOW_HIST>
__device__ __forceinline__ BlockResult decode_lz4_block(const uint8_t *__restrict__ src, uint32_t n,
							uint8_t *o, uint32_t cap, uint32_t hist, int lane)
{
	BlockResult r = {LZ4B200_ST_OK, 0, 0, 0};
	uint32_t ip = 0, op = 0;
	while (ip < n) {
		const uint32_t token = ld_u8<true>(src + ip);
		ip += 1;
		uint32_t lit = token >> 4;
		if (lit == 15 && !read_length_ext(src, ip, n, lit, lane)) {
			r.code = LZ4B200_ST_LIT_EXT_OVERRUN; r.err_pos = op; break;
		}
		if (lit > 0) {
			if (lit > n - ip) {
				// reference: reads past the block, then reports :754 if the nibble /= 0
				r.code = (token & 15) ? LZ4B200_ST_ENDS_AFTER_LITERALS : LZ4B200_ST_LITERAL_OVERRUN;
				r.aux = (token & 15) ? static_cast<int32_t>(token & 15) : static_cast<int32_t>(n - ip);
				r.err_pos = op + lit;   // content-size accounting already saw these bytes
				r.out_len = lit;        // reported in the LITERAL_OVERRUN message
				break;
			}
			if (lit > cap - op) { r.code = LZ4B200_ST_OUTPUT_OVERFLOW; r.err_pos = op + lit; break; }
			warp_copy<true>(o + op, src + ip, lit, lane);
			ip += lit;
			op += lit;
		}
		if (ip >= n) {   // :752-764
			if (token & 15) {
				r.code = LZ4B200_ST_ENDS_AFTER_LITERALS; r.aux = token & 15; r.err_pos = op;
			}
			break;
		}
		if (ip + 1 >= n) { r.code = LZ4B200_ST_OFFSET_TRUNCATED; r.err_pos = op; break; }
		const uint32_t offset = ld_u8<true>(src + ip) | (ld_u8<true>(src + ip + 1) << 8);
		ip += 2;
		if (offset == 0) { r.code = LZ4B200_ST_OFFSET_ZERO; r.err_pos = op; break; }
		uint32_t ml = token & 15;
		if (ml == 15 && !read_length_ext(src, ip, n, ml, lane)) {
			r.code = LZ4B200_ST_MATCH_EXT_OVERRUN; r.err_pos = op; break;
		}
		ml += 4;
		if (offset > op) {
			// reaches before this block: legal only inside the frame's history (:864-874)
			if (hist != 0xffffffffu && offset - op > hist) {
				r.code = LZ4B200_ST_BACKREF_RANGE;
				r.aux = static_cast<int32_t>(hist + op) - static_cast<int32_t>(offset);
				r.err_pos = op;
				break;
			}
			if (!ALLOW_HIST) { r.code = LZ4B200_ST_NEEDS_HISTORY; r.err_pos = op; break; }
		}
		if (ml > cap - op) { r.code = LZ4B200_ST_OUTPUT_OVERFLOW; r.err_pos = op + ml; break; }
		match_copy(o + op, offset, ml, lane);
		op += ml;
	}
	if (r.code == LZ4B200_ST_OK) r.out_len = op;
	return r;
}

// One block, start to finish, by one warp: optional fused block checksum over the payload
// (Check_Checksum, lib/lz4ada.adb:698-707 -- verified before any decoding, :672-676), then
// either the stored copy (:685-695) or the LZ4 decode.
template <bool ALLOW_HIST>
__device__ __forceinline__ void process_block(const uint8_t *__restrict__ src_base, uint8_t *o,
					      const lz4b200_blk_desc &d, uint32_t cap, uint32_t hist,
					      lz4b200_blk_status *st, int lane)
{
	const uint8_t *s = src_base + d.src_off;
	uint32_t computed = 0, declared = 0;
	BlockResult r = {LZ4B200_ST_OK, 0, 0, 0};
	if (d.flags & LZ4B200_BLK_HAS_CHECKSUM) {
		const uint8_t *t = s + d.src_len;
		declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) |
			   (ld_u8<true>(t + 3) << 24);
		computed = quad_xxh32<true, false>(s, d.src_len, lane);
		if (computed != declared) r.code = LZ4B200_ST_BLOCK_CHECKSUM;
	}
	if (r.code == LZ4B200_ST_OK && !(d.flags & LZ4B200_BLK_HASH_ONLY)) {
		if (d.flags & LZ4B200_BLK_STORED) {
			if (d.src_len > cap) {
				r.code = LZ4B200_ST_OUTPUT_OVERFLOW; r.err_pos = d.src_len;
			} else {
				warp_copy<true>(o, s, d.src_len, lane);
				r.out_len = d.src_len;
			}
		} else {
			r = decode_lz4_block<ALLOW_HIST>(s, d.src_len, o, cap, hist, lane);
		}
	}
	if (lane == 0) {
		st->code = r.code;
		st->out_len = r.out_len;
		st->err_pos = r.err_pos;
		st->aux = r.aux;
		st->xxh32_computed = computed;
		st->xxh32_declared = declared;
	}
}

}  // namespace lz4b200
