// shim.cu -- the lz4b200_* half of include/lz4b200.h: device context, memory, transfers and the
// kernel launches K1..K5.  Compiled by nvcc for sm_100a only; everything exported is extern "C".
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "kernels.cuh"
#include "kernels_v2.cuh"
#include "kernels_v3.cuh"
#include "kernels_v4.cuh"
#include "kernels_v5.cuh"
#include "kernels_v6.cuh"
#include "kernels_k6.cuh"
#include "kernels_k7.cuh"
#include "lz4b200.h"

using namespace lz4b200;

// ------------------------------------------------------------------------------------------
// Kernels
// ------------------------------------------------------------------------------------------

constexpr int K1_WARPS = 4;   // warps per CTA, one block per warp

// K1 (+K2): independent blocks, one warp per block.
__global__ void __launch_bounds__(K1_WARPS * 32)
decode_blocks_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
		     const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status)
{
	const int lane = threadIdx.x & 31;
	const uint32_t b = blockIdx.x * K1_WARPS + (threadIdx.x >> 5);
	if (b >= n_blocks) return;
	const lz4b200_blk_desc d = desc[b];
	if (d.flags & LZ4B200_BLK_NOT_K1) return;
	process_block<false>(src, dst + d.dst_off, d, d.dst_cap, d.hist_avail, status + b, lane);
}

// K1 fast path: G blocks per warp (kernels_v2.cuh); falls back to process_block per block.
template <int G>
__global__ void __launch_bounds__(K1_WARPS * 32, 7)
decode_blocks_v2_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
			const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status)
{
	__shared__ SeqDesc sd[K1_WARPS][G * SD_STRIDE];
	__shared__ uint4 tiles[K1_WARPS][(TILE_BYTES + 32) / 16];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const uint32_t first = (blockIdx.x * K1_WARPS + warp) * G;
	if (first >= n_blocks) return;
	decode_group<G>(src, dst, n_blocks, first, desc, status, sd[warp], reinterpret_cast<uint8_t *>(tiles[warp]), lane);
}

// K1 third generation (kernels_v3.cuh): a CTA per block -- eight decode warps with the block's output
// window in shared memory, one warp hashing the payloads of the CTA's blocks.
__global__ void __launch_bounds__(v3::CTA_THREADS, 2)
decode_blocks_v3_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
			const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, uint32_t per_cta,
			unsigned long long *prof)
{
	extern __shared__ __align__(16) uint8_t v3_smem[];
	const uint32_t first = blockIdx.x * per_cta;
	if (first >= n_blocks) return;
	const uint32_t cnt = n_blocks - first < per_cta ? n_blocks - first : per_cta;
	uint32_t computed = 0, declared = 0;
	bool want = false;
	if (threadIdx.x >= v3::NT)
		v3::hash_role(src, first, cnt, desc, status, computed, declared, want);
	else
		v3::decode_role(src, dst, n_blocks, first, cnt, desc, status, v3_smem, prof);
	__syncthreads();
	// Check_Checksum comes before any decoding (lib/lz4ada.adb:672-676): a wrong block checksum beats
	// whatever the decode role reported for the block
	if (threadIdx.x >= v3::NT && want && (threadIdx.x & 3) == 0 && computed != declared) {
		lz4b200_blk_status *st = status + first + ((threadIdx.x & 31) >> 2);
		st->code = LZ4B200_ST_BLOCK_CHECKSUM;
		st->out_len = 0;
		st->err_pos = 0;
		st->aux = 0;
	}
}

// K1 fourth generation (kernels_v4.cuh): a warp per block (G blocks per warp, hashed together and
// decoded in turn), parallel parse and a 4 KiB output ring per warp in shared memory.
__global__ void __launch_bounds__(v4::WARPS * 32, 5)
decode_blocks_v4_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
			const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, uint32_t G)
{
	extern __shared__ __align__(16) uint8_t v4_smem[];
	v4::WarpMem *wms = reinterpret_cast<v4::WarpMem *>(v4_smem);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t first = (blockIdx.x * (blockDim.x >> 5) + warp) * G;   // 1 or v4::WARPS warps per CTA
	if (first >= n_blocks) return;
	v4::decode_group(src, dst, n_blocks, first, G, desc, status, wms[warp], lane);
}

// K1 fifth generation (kernels_v5.cuh): a LANE per block, 32 blocks in lock-step per warp, per-lane
// rings in shared memory (bank-per-lane layout); lanes take blocks from a global counter.
__global__ void __launch_bounds__(v5::WARPS * 32, 8)
decode_blocks_v5_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
			const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, uint32_t *counter)
{
	extern __shared__ __align__(16) uint8_t v5_smem[];
	v5::WarpMem *wms = reinterpret_cast<v5::WarpMem *>(v5_smem);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	v5::decode_lanes(src, dst, n_blocks, desc, status, counter, wms[warp], lane);
}

// LZ4B200_BLK_RING_CAP: the block's dst_cap is the length of the reference caller's Buffer; what the block may
// produce is what is left of that Buffer behind the ring cursor (lib/lz4ada.adb:678-680: the cursor goes back to 0
// at a block start once it has passed the 64 KiB history).  `ring` = cursor before this block, updated in place.
__device__ __forceinline__ uint32_t ring_block_cap(const lz4b200_blk_desc &d, uint32_t &ring)
{
	if (!(d.flags & LZ4B200_BLK_RING_CAP)) return d.dst_cap;
	if (ring >= 65536u) ring = 0;
	return d.dst_cap > ring ? d.dst_cap - ring : 0u;
}

// K1 sixth generation (kernels_v6.cuh): a lane per block again, one piece of a sequence (<= 8 literal + <= 16 match bytes) per
// trip, the parse K pieces ahead of the copy (old match sources requested that far ahead), literals inside the descriptors.
template <uint32_t OWW, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
decode_blocks_v6_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_blocks,
			const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, uint32_t *counter)
{
	// one CTA per SM and launch; the launch picks its number of warps (<= WARPS) from the block count and asks for
	// just their shared memory, so that the launches of several chunks (other streams) can share an SM
	extern __shared__ __align__(16) uint8_t v6_smem[];
	using L = v6::Layout<OWW, K>;
	const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(v6_smem));
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t side = s0 + warp * L::SIDE_WARP;
	const uint32_t in_base = side + lane * v6::IN_STRIDE;
	const uint32_t stage_base = side + L::IN_WARP + lane * v6::STAGE_SLOT;
	const uint32_t rings0 = (s0 + (blockDim.x >> 5) * L::SIDE_WARP + L::OUT_WARP - 1u) & ~(L::OUT_WARP - 1u);
	const uint32_t ring_base = rings0 + warp * L::OUT_WARP + lane * 4u;
	v6::decode_lanes<OWW, K>(src, dst, n_blocks, desc, status, counter, in_base, stage_base, ring_base, static_cast<int>(lane));
}
constexpr int V6A_WARPS = 14, V6A_K = 3;   // 256-byte out rings: 15.5 KB per warp, fourteen warps = 66 304 lanes on 148 SMs
constexpr int V6B_WARPS = 8, V6B_K = 4;    // 512-byte out rings: 24.5 KB per warp

// K4: chains, one warp per chain, blocks in order; the output of a chain is flat, so a match
// simply reads backwards across block boundaries of its frame.
__global__ void __launch_bounds__(K1_WARPS * 32)
decode_linked_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_chains,
		     const lz4b200_chain *__restrict__ chains,
		     const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status)
{
	const int lane = threadIdx.x & 31;
	const uint32_t c = blockIdx.x * K1_WARPS + (threadIdx.x >> 5);
	if (c >= n_chains) return;
	__shared__ SeqDesc sd[K1_WARPS][SD_STRIDE];
	__shared__ uint4 tiles[K1_WARPS][(TILE_BYTES + 32) / 16];
	const lz4b200_chain ch = chains[c];
	uint8_t *out = dst + ch.dst_off;
	uint64_t pos = 0;         // chain-relative output position
	uint64_t frame_start = 0; // chain-relative position where the current frame began
	uint32_t ring = 0;        // Output_Pos of the reference's Buffer (LZ4B200_BLK_RING_CAP blocks)
	bool failed = false;
	for (uint32_t i = 0; i < ch.n_blocks; i++) {
		const uint32_t b = ch.first_block + i;
		if (failed) {
			if (lane == 0) {
				status[b].code = LZ4B200_ST_NOT_RUN;
				status[b].out_len = 0;
				status[b].err_pos = 0;
				status[b].aux = 0;
				status[b].xxh32_computed = 0;
				status[b].xxh32_declared = 0;
			}
			continue;
		}
		const lz4b200_blk_desc d = desc[b];
		if (d.flags & LZ4B200_BLK_FIRST_OF_FRAME) { frame_start = pos; ring = 0; }
		const uint64_t fpos = pos - frame_start;
		const uint32_t hist = fpos > 0xfffffffeull ? 0xffffffffu : static_cast<uint32_t>(fpos);
		const uint64_t room = ch.dst_cap - pos;
		const uint32_t blk_cap = ring_block_cap(d, ring);
		const uint32_t cap = room < blk_cap ? static_cast<uint32_t>(room) : blk_cap;
		// fast path (kernels_v2.cuh) when the frame position fits 32 bits and the block is an
		// ordinary compressed one; the exact routine otherwise and for anything unusual
		bool fast_done = false;
		if (!(d.flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_HASH_ONLY)) && fpos + cap < 0xfff00000ull) {
			uint32_t computed = 0, declared = 0;
			bool sum_ok = true;
			const uint8_t *s = src + d.src_off;
			if (d.flags & LZ4B200_BLK_HAS_CHECKSUM) {
				const uint8_t *t = s + d.src_len;
				declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) |
					   (ld_u8<true>(t + 3) << 24);
				computed = quad_xxh32_prologue(s, d.src_len, lane);
				sum_ok = computed == declared;
			}
			if (sum_ok) {
				uint32_t p32 = static_cast<uint32_t>(fpos);
				if (chain_block_fast(s, d.src_len, out + frame_start, p32, static_cast<uint32_t>(fpos) + cap,
						     sd[threadIdx.x >> 5], reinterpret_cast<uint8_t *>(tiles[threadIdx.x >> 5]), lane)) {
					fast_done = true;
					if (lane == 0) {
						status[b].code = LZ4B200_ST_OK;
						status[b].out_len = p32 - static_cast<uint32_t>(fpos);
						status[b].err_pos = 0;
						status[b].aux = 0;
						status[b].xxh32_computed = computed;
						status[b].xxh32_declared = declared;
					}
				}
			}
		}
		if (!fast_done) process_block<true>(src, out + pos, d, cap, hist, status + b, lane);
		__syncwarp();
		const uint32_t code = status[b].code;      // lane 0 wrote it; visible after __syncwarp
		const uint32_t out_len = status[b].out_len;
		if (code != LZ4B200_ST_OK) failed = true;
		else { pos += out_len; ring += out_len; }
	}
}

// K4 pipelined: one CTA per chain -- warp 0 walks the token chains of the chain's blocks and
// publishes batches of 32 sequences into a shared-memory ring; warps 1..7 copy batches round-robin.
// A big block (4 MiB = ~400 K sequences) or a linked frame is serial in its *parse*; this overlaps
// the parse with the copies and the copies with each other (ordered through BatchGate), instead of
// alternating both in a single warp.
constexpr int PIPE_WARPS = 8;          // 1 parser + 7 copiers (measured: 3 copiers 38.8 ms, 7 copiers 34.2 ms for 256 x 4 MiB text)
constexpr int PIPE_COPIERS = PIPE_WARPS - 1;
constexpr int PIPE_SLOTS = 128;         // a whole parse window (<= 86 batches) must fit beside work in flight
constexpr uint32_t SPEC_SEG = 256;       // compressed bytes per speculative segment (one lane each)
constexpr uint32_t SPEC_VIS = 24;        // token positions a lane publishes for its neighbour to merge with

struct PipeShared {
	SeqDesc sd[PIPE_SLOTS][32];
	uint32_t count[PIPE_SLOTS];
	uint32_t out_start[PIPE_SLOTS];   // frame-relative output position of the batch
	uint32_t cap_abs[PIPE_SLOTS];
	uint32_t frame_base_lo[PIPE_SLOTS], frame_base_hi[PIPE_SLOTS];   // chain-relative start of the batch's frame
	uint32_t src_lo[PIPE_SLOTS], src_hi[PIPE_SLOTS];                 // block payload offset in src
	uint32_t blk[PIPE_SLOTS];
	uint32_t done_flag[PIPE_SLOTS];
	uint16_t vis[32][SPEC_VIS];       // speculative parse: first token positions of each lane's walk (window-relative)
	uint32_t produced, done_upto, fail, end_batch, fail_block;
};

__device__ __forceinline__ uint32_t vload(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void vstore(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// K4 / K6 parser role (one warp): walks the token chains of the chain's blocks -- a speculative parallel parse, 32
// lanes x SPEC_SEG-byte segments, merge-verified, with a serial fallback -- and publishes batches of 32 sequences
// into the shared-memory ring of `ps`.
__device__ __forceinline__ void pipe_parser_role(PipeShared &ps, const lz4b200_chain &ch, const uint8_t *__restrict__ src, uint8_t *out,
						 const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, int lane)
{
	uint64_t pos = 0, frame_start = 0;   // chain-relative
	uint32_t ring = 0;                   // Output_Pos of the reference's Buffer (LZ4B200_BLK_RING_CAP blocks)
	uint32_t k = 0;                      // batches published
	bool stop = false;
	for (uint32_t i = 0; i < ch.n_blocks && !stop; i++) {
		const uint32_t b = ch.first_block + i;
		const lz4b200_blk_desc d = desc[b];
		if (d.flags & LZ4B200_BLK_FIRST_OF_FRAME) { frame_start = pos; ring = 0; }
		const uint64_t fpos0 = pos - frame_start;
		const uint64_t room = ch.dst_cap - pos;
		const uint32_t blk_cap = ring_block_cap(d, ring);
		const uint32_t cap = room < blk_cap ? static_cast<uint32_t>(room) : blk_cap;
		const uint8_t *s = src + d.src_off;
		const bool stored = (d.flags & LZ4B200_BLK_STORED) != 0;
		const bool ordinary = !(d.flags & LZ4B200_BLK_HASH_ONLY) && fpos0 + cap < 0xfff00000ull &&
				      !(stored && d.src_len > cap);
		uint32_t computed = 0, declared = 0;
		bool okay = ordinary;
		if (okay && (d.flags & LZ4B200_BLK_HAS_CHECKSUM)) {
			const uint8_t *t = s + d.src_len;
			declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
			computed = quad_xxh32_prologue(s, d.src_len, lane);
			okay = computed == declared;
		}
		uint32_t fpos = static_cast<uint32_t>(fpos0);
		if (okay && stored) {
			// stored block (lib/lz4ada.adb:685-695): no dependencies, the parser warp copies it itself
			warp_copy<true>(out + pos, s, d.src_len, lane);
			fpos += d.src_len;
		} else if (okay) {
			uint32_t ip = 0;
			const uint32_t n = d.src_len;
			uint32_t spec_backoff = 0;
			while (ip < n && okay) {
				// ---- speculative parallel parse of the next window (32 segments, one lane each) ----
				if (spec_backoff == 0 && n - ip >= 4 * SPEC_SEG) {
					const uint32_t wbase = ip;
					const uint32_t seg_lo = wbase + lane * SPEC_SEG;
					const uint32_t seg_hi = seg_lo + SPEC_SEG < n ? seg_lo + SPEC_SEG : n;
					// pass 1: walk from a guessed token start (lane 0: the true one), publish the first positions
					uint32_t p = seg_lo, cntv = 0;
					bool dead = seg_lo >= n;
					while (!dead && p < seg_hi) {
						uint32_t lp, lit, ml, nx;
						if (!parse_token(s, n, p, lp, lit, ml, nx)) { dead = true; break; }
						if (cntv < SPEC_VIS) ps.vis[lane][cntv] = static_cast<uint16_t>(p - wbase);
						cntv++;
						p = nx;
					}
					const uint32_t exit_pos = p;
					const uint32_t nvis = cntv < SPEC_VIS ? cntv : SPEC_VIS;
					__syncwarp();
					// verification: lane j walks on from its exit until it stands on a position lane j + 1 also
					// visited -- from there on both parse identically, so lane j + 1 is in sync from that token
					const uint32_t nvis_next = __shfl_down_sync(FULL_MASK, nvis, 1);
					const uint32_t lo_next = seg_lo + SPEC_SEG;
					const bool next_exists = lane < 31 && lo_next < n;
					uint32_t bnext = 0xffffffffu;
					bool merged = false;
					if (!dead) {
						if (!next_exists || exit_pos >= n) {
							merged = true;
							bnext = exit_pos;
						} else {
							uint32_t q = exit_pos, kk = 0;
							for (uint32_t step = 0; step < 2 * SPEC_VIS; step++) {
								while (kk < nvis_next && wbase + ps.vis[lane + 1][kk] < q) kk++;
								if (kk >= nvis_next) break;
								if (wbase + ps.vis[lane + 1][kk] == q) { merged = true; bnext = q; break; }
								uint32_t lp, lit, ml, nx;
								if (q >= lo_next + SPEC_SEG || !parse_token(s, n, q, lp, lit, ml, nx)) break;
								q = nx;
							}
						}
					}
					const uint32_t last_lane = (n - 1 - wbase) / SPEC_SEG < 31 ? (n - 1 - wbase) / SPEC_SEG : 31;
					const uint32_t need_mask = last_lane >= 31 ? 0xffffffffu : ((2u << last_lane) - 1u);
					const uint32_t ok_mask = __ballot_sync(FULL_MASK, merged);
					bool spec_ok = (ok_mask & need_mask) == need_mask;
					uint32_t b_start = __shfl_up_sync(FULL_MASK, bnext, 1);
					if (lane == 0) b_start = wbase;
					const bool mine = static_cast<uint32_t>(lane) <= last_lane;
					// pass 2: count the true tokens of [b_start, bnext)
					uint32_t cnt_t = 0, out_t = 0;
					if (spec_ok && mine) {
						uint32_t q = b_start;
						while (q < bnext) {
							uint32_t lp, lit, ml, nx;
							if (!parse_token(s, n, q, lp, lit, ml, nx)) { cnt_t = 0xffffffffu; break; }
							cnt_t++;
							out_t += lit + ml;
							q = nx;
						}
						if (cnt_t != 0xffffffffu && q != bnext) cnt_t = 0xffffffffu;
					}
					if (__any_sync(FULL_MASK, cnt_t == 0xffffffffu)) spec_ok = false;
					if (spec_ok) {
						uint32_t icnt = cnt_t, iout = out_t;
#pragma unroll
						for (int sh = 1; sh < 32; sh <<= 1) {
							const uint32_t a = __shfl_up_sync(FULL_MASK, icnt, sh), bsum = __shfl_up_sync(FULL_MASK, iout, sh);
							if (lane >= sh) { icnt += a; iout += bsum; }
						}
						const uint32_t tot_cnt = __shfl_sync(FULL_MASK, icnt, 31), tot_out = __shfl_sync(FULL_MASK, iout, 31);
						const uint32_t nb = (tot_cnt + 31) / 32;
						const uint32_t used = fpos - static_cast<uint32_t>(fpos0);
						if (tot_out > cap - used || nb > PIPE_SLOTS - 32 || tot_cnt == 0) {
							spec_ok = false;   // capacity: the exact path reports it; nb: never with 256-byte segments
						} else {
							if (lane == 0) {
								while (k + nb - vload(&ps.done_upto) > PIPE_SLOTS) __nanosleep(40);
							}
							__syncwarp();
							if (vload(&ps.fail) != 0) { okay = false; break; }
							// pass 3: emit descriptors straight into the ring at their global sequence index
							if (mine) {
								uint32_t idx = icnt - cnt_t, opos = fpos + (iout - out_t), q = b_start;
								while (q < bnext) {
									uint32_t lp, lit, ml, nx;
									parse_token(s, n, q, lp, lit, ml, nx);
									const uint32_t sl = (k + (idx >> 5)) % PIPE_SLOTS;
									*reinterpret_cast<uint2 *>(&ps.sd[sl][idx & 31]) = make_uint2(lp, lit | (ml << 16));
									if ((idx & 31) == 0) {
										ps.count[sl] = tot_cnt - idx < 32 ? tot_cnt - idx : 32;
										ps.out_start[sl] = opos;
										ps.cap_abs[sl] = static_cast<uint32_t>(fpos0) + cap;
										ps.frame_base_lo[sl] = static_cast<uint32_t>(frame_start);
										ps.frame_base_hi[sl] = static_cast<uint32_t>(frame_start >> 32);
										ps.src_lo[sl] = static_cast<uint32_t>(d.src_off);
										ps.src_hi[sl] = static_cast<uint32_t>(d.src_off >> 32);
										ps.blk[sl] = b;
									}
									idx++;
									opos += lit + ml;
									q = nx;
								}
							}
							__syncwarp();
							__threadfence_block();
							if (lane == 0) vstore(&ps.produced, k + nb);
							k += nb;
							fpos += tot_out;
							ip = __shfl_sync(FULL_MASK, bnext, last_lane);
							continue;
						}
					}
					spec_backoff = 8;   // segments did not re-synchronise here (long literal runs): go serial for a while
				}
				if (spec_backoff) spec_backoff--;
				// ---- serial: wait for a free slot (lane 0 polls), then parse up to 32 sequences into it ----
				uint32_t cnt = 0, total = 0;
				bool fb = false;
				// the idle lanes pull the next kilobyte of the compressed stream towards L1 so that the
				// token chain walked by lane 0 does not pay a DRAM round trip every 128 bytes
				if (lane >= 1 && lane <= 8 && ip + 128u * lane < n)
					asm volatile("prefetch.global.L1 [%0];" ::"l"(s + ip + 128u * lane));
				if (lane == 0) {
					// every published batch is eventually marked done (copied or skipped), so this ends
					while (k - vload(&ps.done_upto) >= PIPE_SLOTS) __nanosleep(40);
					SeqDesc *my = ps.sd[k % PIPE_SLOTS];
					if (vload(&ps.fail) != 0) fb = true;   // a copier gave up: stop feeding
					while (!fb && cnt < 32 && ip < n) {
						const uint32_t t = ld_u8<true>(s + ip);
						uint32_t lit = t >> 4, ml = t & 15, p = ip + 1, nxt;
						if (lit == 15 || ml == 15) {
							if (!parse_extended(s, n, p, lit, ml, nxt)) { fb = true; break; }
						} else {
							const uint32_t q = p + lit;
							if (q + 2 <= n) { ml += 4; nxt = q + 2; }
							else if (q == n && ml == 0) { nxt = n; }
							else { fb = true; break; }
						}
						*reinterpret_cast<uint2 *>(my + cnt) = make_uint2(p, lit | (ml << 16));
						cnt++;
						total += lit + ml;
						ip = nxt;
					}
					if (total > cap - (fpos - static_cast<uint32_t>(fpos0))) fb = true;   // exact path reports it
					if (!fb && cnt) {
						const uint32_t sl = k % PIPE_SLOTS;
						ps.count[sl] = cnt;
						ps.out_start[sl] = fpos;
						ps.cap_abs[sl] = static_cast<uint32_t>(fpos0) + cap;
						ps.frame_base_lo[sl] = static_cast<uint32_t>(frame_start);
						ps.frame_base_hi[sl] = static_cast<uint32_t>(frame_start >> 32);
						ps.src_lo[sl] = static_cast<uint32_t>(d.src_off);
						ps.src_hi[sl] = static_cast<uint32_t>(d.src_off >> 32);
						ps.blk[sl] = b;
						__threadfence_block();
						vstore(&ps.produced, k + 1);
					}
				}
				fb = __shfl_sync(FULL_MASK, fb ? 1 : 0, 0) != 0;
				cnt = __shfl_sync(FULL_MASK, cnt, 0);
				total = __shfl_sync(FULL_MASK, total, 0);
				ip = __shfl_sync(FULL_MASK, ip, 0);
				if (fb || vload(&ps.fail) != 0) { okay = false; break; }
				if (cnt) { k++; fpos += total; }
			}
		}
		if (okay) {
			// provisional: stands unless a copier gives up on one of this block's batches
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
			if (lane == 0 && vload(&ps.fail_block) == 0xffffffffu) atomicMin(&ps.fail_block, i);
			stop = true;
		}
	}
	if (lane == 0) {
		__threadfence_block();
		vstore(&ps.end_batch, k);
	}
}

// The exact routine from the first block the fast path gave up on (one warp, after the pipeline has drained).
__device__ __forceinline__ void pipe_exact_tail(PipeShared &ps, const lz4b200_chain &ch, const uint8_t *__restrict__ src, uint8_t *out,
						const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, int lane)
{
	const int warp = 0;
	const uint32_t fbk = ps.fail_block;
	if (fbk != 0xffffffffu && warp == 0) {
		uint64_t pos = 0, frame_start = 0;
		uint32_t ring = 0;
		bool failed = false;
		for (uint32_t i = 0; i < ch.n_blocks; i++) {
			const uint32_t b = ch.first_block + i;
			const lz4b200_blk_desc d = desc[b];
			if (d.flags & LZ4B200_BLK_FIRST_OF_FRAME) { frame_start = pos; ring = 0; }
			if (i < fbk) {   // finished by the pipeline
				ring_block_cap(d, ring);   // (the cursor's wrap at this block's start)
				pos += status[b].out_len;
				ring += status[b].out_len;
				continue;
			}
			if (failed) {
				if (lane == 0) {
					status[b].code = LZ4B200_ST_NOT_RUN;
					status[b].out_len = 0; status[b].err_pos = 0; status[b].aux = 0;
					status[b].xxh32_computed = 0; status[b].xxh32_declared = 0;
				}
				continue;
			}
			const uint64_t fpos = pos - frame_start;
			const uint64_t room = ch.dst_cap - pos;
			const uint32_t blk_cap = ring_block_cap(d, ring);
			const uint32_t cap = room < blk_cap ? static_cast<uint32_t>(room) : blk_cap;
			if (d.flags & LZ4B200_BLK_SOLO) {
				// a chain of one block taken out of an independent frame: independent semantics
				process_block<false>(src, out + pos, d, cap, d.hist_avail, status + b, lane);
			} else {
				const uint32_t hist = fpos > 0xfffffffeull ? 0xffffffffu : static_cast<uint32_t>(fpos);
				process_block<true>(src, out + pos, d, cap, hist, status + b, lane);
			}
			__syncwarp();
			if (status[b].code != LZ4B200_ST_OK) failed = true;
			else { pos += status[b].out_len; ring += status[b].out_len; }
		}
	}
}

__global__ void __launch_bounds__(PIPE_WARPS * 32)
decode_chain_pipe_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_chains,
			 const lz4b200_chain *__restrict__ chains, const lz4b200_blk_desc *__restrict__ desc,
			 lz4b200_blk_status *status)
{
	__shared__ PipeShared ps;
	__shared__ uint4 tiles[PIPE_WARPS][(TILE_BYTES + 32) / 16];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t c = blockIdx.x;
	if (c >= n_chains) return;
	const lz4b200_chain ch = chains[c];
	uint8_t *out = dst + ch.dst_off;
	if (threadIdx.x == 0) {
		ps.produced = 0; ps.done_upto = 0; ps.fail = 0; ps.end_batch = 0xffffffffu; ps.fail_block = 0xffffffffu;
	}
	if (threadIdx.x < PIPE_SLOTS) ps.done_flag[threadIdx.x] = 0;
	__syncthreads();

	if (warp == 0) {
		pipe_parser_role(ps, ch, src, out, desc, status, lane);
	} else {
		// ---------------- copiers ----------------
		uint8_t *tile = reinterpret_cast<uint8_t *>(tiles[warp]);
		for (uint32_t k = warp - 1;; k += PIPE_COPIERS) {
			uint32_t go = 0;
			if (lane == 0) {
				for (;;) {
					if (vload(&ps.produced) > k) { go = 1; break; }
					if (vload(&ps.end_batch) <= k) { go = 0; break; }
					__nanosleep(40);
				}
			}
			go = __shfl_sync(FULL_MASK, go, 0);
			if (!go) break;
			__threadfence_block();
			const uint32_t sl = k % PIPE_SLOTS;
			const uint32_t cnt = ps.count[sl], ostart = ps.out_start[sl], cap_abs = ps.cap_abs[sl], blk = ps.blk[sl];
			const uint64_t fbase = ps.frame_base_lo[sl] | (static_cast<uint64_t>(ps.frame_base_hi[sl]) << 32);
			const uint64_t soff = ps.src_lo[sl] | (static_cast<uint64_t>(ps.src_hi[sl]) << 32);
			// Batches of blocks before the first failing block must still be completed (they are final
			// output); batches from the failing block on are redone by the exact routine: skip them.
			const bool skip = blk - ch.first_block >= vload(&ps.fail_block);
			bool okay = true;
			if (!skip) {
				BatchGate gate;
				gate.done_upto = &ps.done_upto;
				gate.out_start = ps.out_start;
				gate.base_lo = ps.frame_base_lo;
				gate.base_hi = ps.frame_base_hi;
				gate.my_base_lo = ps.frame_base_lo[sl];
				gate.my_base_hi = ps.frame_base_hi[sl];
				gate.fail = &ps.fail;
				gate.slots = PIPE_SLOTS;
				gate.my_batch = k;
				uint32_t total = 0;
				okay = copy_batch<true>(src + soff, out + fbase, ostart, cap_abs, ps.sd[sl], cnt, lane, tile, total, &gate);
			}
			__syncwarp();
			if (lane == 0) {
				if (!okay) {
					// which block failed first decides where the exact routine restarts
					atomicMin(&ps.fail_block, blk - ch.first_block);
					vstore(&ps.fail, 1);
				}
				__threadfence_block();
				vstore(&ps.done_flag[sl], k + 1);
				// advance the in-order frontier
				uint32_t dn = vload(&ps.done_upto);
				while (vload(&ps.done_flag[dn % PIPE_SLOTS]) == dn + 1) dn++;
				atomicMax(&ps.done_upto, dn);
			}
			__syncwarp();
		}
	}
	__syncthreads();
	if (warp == 0) pipe_exact_tail(ps, ch, src, out, desc, status, lane);
}

// K6: one CTA of two warps per chain -- warp 0 parses in steps of 1 KiB of staged compressed bytes (pointer doubling
// over a next-token table), warp 1 copies the batches in order, each in dependency rounds on a 64 KiB shared-memory
// window of the chain's output (kernels_k6.cuh).
constexpr uint32_t K6_SMEM = k6::WIN + ((sizeof(k6::Shared) + 15) & ~15u) + TILE_BYTES + 32 + 32 * sizeof(SeqDesc) + 64;

// The exact routine from the first block the fast path gave up on (one warp, after the pipeline has drained).
__device__ __forceinline__ void chain_exact_tail(uint32_t fbk, const lz4b200_chain &ch, const uint8_t *__restrict__ src, uint8_t *out,
						 const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, int lane)
{
	if (fbk == 0xffffffffu) return;
	uint64_t pos = 0, frame_start = 0;
	uint32_t ring = 0;
	bool failed = false;
	for (uint32_t i = 0; i < ch.n_blocks; i++) {
		const uint32_t b = ch.first_block + i;
		const lz4b200_blk_desc d = desc[b];
		if (d.flags & LZ4B200_BLK_FIRST_OF_FRAME) { frame_start = pos; ring = 0; }
		if (i < fbk) {   // finished by the pipeline
			ring_block_cap(d, ring);   // (the cursor's wrap at this block's start)
			pos += status[b].out_len;
			ring += status[b].out_len;
			continue;
		}
		if (failed) {
			if (lane == 0) {
				status[b].code = LZ4B200_ST_NOT_RUN;
				status[b].out_len = 0; status[b].err_pos = 0; status[b].aux = 0;
				status[b].xxh32_computed = 0; status[b].xxh32_declared = 0;
			}
			continue;
		}
		const uint64_t fpos = pos - frame_start;
		const uint64_t room = ch.dst_cap - pos;
		const uint32_t blk_cap = ring_block_cap(d, ring);
		const uint32_t cap = room < blk_cap ? static_cast<uint32_t>(room) : blk_cap;
		if (d.flags & LZ4B200_BLK_SOLO) {
			// a chain of one block taken out of an independent frame: independent semantics
			process_block<false>(src, out + pos, d, cap, d.hist_avail, status + b, lane);
		} else {
			const uint32_t hist = fpos > 0xfffffffeull ? 0xffffffffu : static_cast<uint32_t>(fpos);
			process_block<true>(src, out + pos, d, cap, hist, status + b, lane);
		}
		__syncwarp();
		if (status[b].code != LZ4B200_ST_OK) failed = true;
		else { pos += status[b].out_len; ring += status[b].out_len; }
	}
}

__global__ void __launch_bounds__(64, 2)
decode_chain_k6_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_chains,
		       const lz4b200_chain *__restrict__ chains, const lz4b200_blk_desc *__restrict__ desc,
		       lz4b200_blk_status *status, uint32_t dbg)
{
	extern __shared__ __align__(16) uint8_t k6_smem[];
	k6::Shared &sh = *reinterpret_cast<k6::Shared *>(k6_smem + k6::WIN);
	uint8_t *tile = k6_smem + k6::WIN + ((sizeof(k6::Shared) + 15) & ~size_t(15));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t c = blockIdx.x;
	if (c >= n_chains) return;
	const lz4b200_chain ch = chains[c];
	uint8_t *out = dst + ch.dst_off;
	if (threadIdx.x == 0) {
		sh.produced = 0; sh.done_upto = 0; sh.fail = 0; sh.end_batch = 0xffffffffu; sh.fail_block = 0xffffffffu;
		sh.copier_in = 0; sh.copier_blk = 0;
	}
	__syncthreads();
	if (warp == 0) k6::parser_role(sh, ch, src, desc, status, lane);
	else k6::copier_role(sh, static_cast<uint32_t>(__cvta_generic_to_shared(k6_smem)), tile, ch, src, out, lane, dbg);
	__syncthreads();
	if (warp == 0) chain_exact_tail(sh.fail_block, ch, src, out, desc, status, lane);
}

// K7: one CTA per chain, no serial walk (kernels_k7.cuh): a speculative parse by every thread, then the output built
// 16 KiB at a time in shared memory with match bytes as pointers that are resolved by pointer jumping.
constexpr uint32_t K7_SMEM = sizeof(k7::Shared);
__global__ void __launch_bounds__(k7::T, k7::T == 512 ? 2 : 3)
decode_chain_k7_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_chains,
		       const lz4b200_chain *__restrict__ chains, const lz4b200_blk_desc *__restrict__ desc,
		       lz4b200_blk_status *status)
{
	extern __shared__ __align__(16) uint8_t k7_smem[];
	k7::Shared &sh = *reinterpret_cast<k7::Shared *>(k7_smem);
	const uint32_t c = blockIdx.x;
	if (c >= n_chains) return;
	const lz4b200_chain ch = chains[c];
	uint8_t *out = dst + ch.dst_off;
	const uint32_t fb = k7::run_chain(sh, ch, src, out, desc, status);
	if (threadIdx.x < 32) chain_exact_tail(fb, ch, src, out, desc, status, threadIdx.x & 31);
}

// ... and in front of it: the block checksums of every chained block (Check_Checksum, lib/lz4ada.adb:698-707), a quad
// per block and all blocks of the launch at once instead of one serial hash after the other inside the chain's CTA.
// Results go to the status entries (xxh32_computed / xxh32_declared), where the chain kernel reads them.
__global__ void __launch_bounds__(128)
chain_block_hashes_kernel(const uint8_t *__restrict__ src, uint32_t n_chains, const lz4b200_chain *__restrict__ chains,
			  const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status)
{
	const uint32_t c = blockIdx.x;
	if (c >= n_chains) return;
	const lz4b200_chain ch = chains[c];
	const int lane = threadIdx.x & 31;
	const uint32_t quad = threadIdx.x >> 2;
	for (uint32_t i0 = 0; i0 < ch.n_blocks; i0 += 32) {
		const uint32_t i = i0 + quad;
		const bool have = i < ch.n_blocks;
		const uint32_t b = ch.first_block + (have ? i : 0u);
		const lz4b200_blk_desc d = desc[b];
		const bool want = have && (d.flags & LZ4B200_BLK_HAS_CHECKSUM) && !(d.flags & LZ4B200_BLK_HASH_ONLY);
		if (!__any_sync(FULL_MASK, want)) continue;
		const uint8_t *s = src + d.src_off;
		const uint32_t h = quad_xxh32_prologue(s, want ? d.src_len : 0u, lane);
		if (want && (lane & 3) == 0) {
			const uint8_t *t = s + d.src_len;
			status[b].xxh32_computed = h;
			status[b].xxh32_declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
		}
	}
}

// One block against a device-resident history window (single-block path under Update).
// The K1 v4 block decoder (parallel parse, shared-memory ring) with the window in front of the cursor as
// history; the exact routine takes over for anything unusual.
__global__ void __launch_bounds__(32)
stream_block_kernel(const uint8_t *__restrict__ src, uint8_t *win, const lz4b200_blk_desc *desc,
		    lz4b200_blk_status *status)
{
	__shared__ v4::WarpMem wm;
	const int lane = threadIdx.x & 31;
	const lz4b200_blk_desc d = *desc;
	if (!(d.flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_HASH_ONLY)) && d.dst_off < 0x70000000ull && d.dst_cap < 0x08000000u) {
		uint32_t computed = 0, declared = 0;
		bool sum_ok = true;
		const uint8_t *s = src + d.src_off;
		if (d.flags & LZ4B200_BLK_HAS_CHECKSUM) {   // verified before any decoding, lib/lz4ada.adb:672-676
			const uint8_t *t = s + d.src_len;
			declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
			computed = quad_xxh32_prologue(s, lane < 4 ? d.src_len : 0u, lane);
			computed = __shfl_sync(FULL_MASK, computed, 0);
			sum_ok = computed == declared;
		}
		if (sum_ok) {
			// history the reference would accept: the frame's bytes in front of the cursor, at most what the
			// window holds (hist_avail = 0xffffffff once its 64 KiB ring has wrapped)
			const uint32_t before = static_cast<uint32_t>(d.dst_off);
			const uint32_t hist = d.hist_avail < before ? d.hist_avail : before;
			uint32_t produced = 0;
			if (v4::decode_block(s, d.src_len, win + d.dst_off - hist, d.dst_cap, wm, lane, produced, hist)) {
				if (lane == 0) {
					status->code = LZ4B200_ST_OK;
					status->out_len = produced;
					status->err_pos = 0;
					status->aux = 0;
					status->xxh32_computed = computed;
					status->xxh32_declared = declared;
				}
				return;
			}
		}
		__syncwarp();
	}
	process_block<true>(src, win + d.dst_off, d, d.dst_cap, d.hist_avail, status, lane);
}

// ... and a big block under Update: the same history window, decoded by a whole CTA with the chain kernel's machinery
// (kernels_k7.cuh) instead of one warp; the exact routine takes over for anything unusual, as above.
__global__ void __launch_bounds__(k7::T, k7::T == 512 ? 2 : 3)
stream_block_k7_kernel(const uint8_t *__restrict__ src, uint8_t *win, const lz4b200_blk_desc *desc, lz4b200_blk_status *status)
{
	extern __shared__ __align__(16) uint8_t k7_smem[];
	k7::Shared &sh = *reinterpret_cast<k7::Shared *>(k7_smem);
	const int lane = threadIdx.x & 31;
	const lz4b200_blk_desc d = *desc;
	uint32_t fb = 0;
	if (!(d.flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_HASH_ONLY)) && d.dst_off < 0x70000000ull && d.dst_cap < 0x08000000u) {
		if ((d.flags & LZ4B200_BLK_HAS_CHECKSUM) && threadIdx.x < 32) {   // verified before any decoding, lib/lz4ada.adb:672-676
			const uint8_t *s = src + d.src_off;
			const uint8_t *t = s + d.src_len;
			uint32_t computed = quad_xxh32_prologue(s, lane < 4 ? d.src_len : 0u, lane);
			computed = __shfl_sync(FULL_MASK, computed, 0);
			if (lane == 0) {
				status->xxh32_computed = computed;
				status->xxh32_declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
			}
		}
		__syncthreads();
		// history the reference would accept: the frame's bytes in front of the cursor, at most what the window holds
		const uint32_t before = static_cast<uint32_t>(d.dst_off);
		const uint32_t hist = d.hist_avail < before ? d.hist_avail : before;
		lz4b200_chain ch;
		ch.first_block = 0;
		ch.n_blocks = 1;
		ch.dst_off = 0;
		ch.dst_cap = static_cast<uint64_t>(hist) + d.dst_cap;
		fb = k7::run_chain(sh, ch, src, win + d.dst_off - hist, desc, status, hist);
		if (fb == 0xffffffffu) return;
	}
	__syncthreads();
	if (threadIdx.x < 32) process_block<true>(src, win + d.dst_off, d, d.dst_cap, d.hist_avail, status, lane);
}

// K3: one XXH32 chain per byte range, eight ranges per warp (a quad each).
__global__ void __launch_bounds__(128)
xxh32_spans_kernel(const uint8_t *__restrict__ data, uint32_t n,
		   const lz4b200_hash_span *__restrict__ spans, uint32_t *out)
{
	const int lane = threadIdx.x & 31;
	const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const uint32_t idx = warp * 8 + (lane >> 2);
	const bool valid = idx < n;
	uint64_t off = 0, len = 0;
	if (valid) {
		off = spans[idx].off;
		len = spans[idx].len;
	}
	extern __shared__ uint4 xxh_rings[];
	uint8_t *ring = reinterpret_cast<uint8_t *>(xxh_rings) + (threadIdx.x >> 5) * (8 * XXH_RING_STRIDE);
	const uint32_t h = quad_xxh32_stream(data + off, len, ring, lane);
	if (valid && (lane & 3) == 0) out[idx] = h;
}

// K3 (batch): content checksum per frame, length summed from the block statuses on the device.
template <uint32_t GROUP_BYTES>
__global__ void __launch_bounds__(128)
xxh32_frames_kernel(const uint8_t *__restrict__ dst, uint32_t n_frames,
		    const lz4b200_frame_blocks *__restrict__ frames,
		    const lz4b200_blk_desc *__restrict__ desc, const lz4b200_blk_status *__restrict__ status,
		    uint32_t *digest, uint32_t *valid)
{
	const int lane = threadIdx.x & 31;
	const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const uint32_t f = warp * 8 + (lane >> 2);
	const bool have = f < n_frames;
	uint64_t base = 0, len = 0;
	bool okay = have;
	if (have) {
		const lz4b200_frame_blocks fb = frames[f];
		if (fb.n_blocks) base = desc[fb.first_block].dst_off;
		for (uint32_t i = 0; i < fb.n_blocks; i++) {
			const uint32_t b = fb.first_block + i;
			const uint32_t fl = desc[b].flags;
			const bool placed_by_table = !(fl & LZ4B200_BLK_CHAINED) || (fl & LZ4B200_BLK_SOLO);
			if (status[b].code != LZ4B200_ST_OK || (placed_by_table && desc[b].dst_off != base + len)) {
				okay = false;
				break;
			}
			len += status[b].out_len;
		}
	}
	if (!okay) len = 0;
	extern __shared__ uint4 xxh_rings[];
	uint8_t *ring = reinterpret_cast<uint8_t *>(xxh_rings) + (threadIdx.x >> 5) * (8 * (XXH_GROUPS * GROUP_BYTES + XXH_RING_PAD));
	const uint32_t h = quad_xxh32_stream_t<GROUP_BYTES>(dst + base, len, ring, lane);
	if (have && (lane & 3) == 0) {
		digest[f] = h;
		valid[f] = okay ? 1u : 0u;
	}
}

// Streaming content checksum for the single-block path under Update: the XXH32 state lives in
// device memory between blocks (XXHash32.Update, lib/lz4ada.adb:942-977).
// ring: XXH_HUGE_RING_STRIDE bytes of shared memory -- the stripes run through the cp.async ring of K3 (quad 0 of the warp:
// 16 KiB in flight, 2 GB/s) instead of register-staged loads (0.6 GB/s, which made the content checksum the slowest
// part of the read-ahead path of Update).
__device__ __forceinline__ void xxh_stream_update(XxhState *st, const uint8_t *__restrict__ data, uint32_t len, uint32_t *digest, int lane,
						   uint8_t *ring)
{
	const int sub = lane & 3;
	// (volatile: in the list kernel the state was written by other lanes of this warp a moment ago)
	uint32_t acc = *reinterpret_cast<volatile uint32_t *>(&st->acc[sub]);
	uint32_t buf_size = *reinterpret_cast<volatile uint32_t *>(&st->buf_size);
	const uint64_t total = *reinterpret_cast<volatile uint64_t *>(&st->total_len) + len;
	uint32_t pos = 0;
	__syncwarp();
	if (buf_size > 0 && len > 0) {
		const uint32_t need = 16 - buf_size;
		const uint32_t take = need < len ? need : len;
		if (lane < take) st->buf[buf_size + lane] = data[lane];
		__syncwarp();
		buf_size += take;
		pos = take;
		if (buf_size == 16) {
			const volatile uint8_t *bp = st->buf + 4 * sub;
			const uint32_t x = bp[0] | (bp[1] << 8) | (bp[2] << 16) | (bp[3] << 24);
			acc = xxh_round(acc, x);
			buf_size = 0;
		}
		__syncwarp();
	}
	const uint32_t nstripes = (len - pos) >> 4;
	{
		const uint32_t mine = quad_ring_stripes_t<XXH_HUGE_GROUP_BYTES>(data + pos, lane < 4 ? nstripes : 0u, acc, ring, lane);
		acc = __shfl_sync(FULL_MASK, mine, sub);   // the chain ran in lanes 0 .. 3
	}
	pos += nstripes << 4;
	const uint32_t rem = len - pos;
	if (rem > 0) {   // only reachable with buf_size == 0
		if (lane < rem) st->buf[lane] = data[pos + lane];
		buf_size = rem;
	}
	__syncwarp();
	if (lane < 4) st->acc[lane] = acc;
	if (lane == 0) {
		st->buf_size = buf_size;
		st->total_len = total;
	}
	if (digest != nullptr) {
		const uint32_t a0 = __shfl_sync(FULL_MASK, acc, 0), a1 = __shfl_sync(FULL_MASK, acc, 1);
		const uint32_t a2 = __shfl_sync(FULL_MASK, acc, 2), a3 = __shfl_sync(FULL_MASK, acc, 3);
		const uint32_t h = xxh_finish<false>(a0, a1, a2, a3, total, st->buf, buf_size);
		if (lane == 0) *digest = h;
	}
	__syncwarp();
}

__global__ void __launch_bounds__(32)
xxh32_stream_kernel(XxhState *st, const uint8_t *__restrict__ data, uint32_t len, uint32_t *digest,
		    const lz4b200_blk_status *cond = nullptr)
{
	// cond: the status of the block decoded just before on the same stream -- hash what it produced, or nothing if
	// it failed (the host has not seen the status yet: no round trip between decode and hash)
	if (cond) {
		if (cond->code != LZ4B200_ST_OK) return;
		len = cond->out_len;
	}
	__shared__ __align__(16) uint8_t ring[XXH_HUGE_RING_STRIDE];
	xxh_stream_update(st, data, len, digest, threadIdx.x & 31, ring);
}

// ... the same over a list of pieces in stream order (the blocks a read-ahead of Update has served): one launch for
// up to 255 blocks instead of one per block.  The list travels as a kernel parameter.
struct PieceList {
	uint32_t n;
	uint32_t off[255];
	uint32_t len[255];
};
__global__ void __launch_bounds__(32)
xxh32_stream_list_kernel(XxhState *st, const uint8_t *__restrict__ base, const PieceList pl)
{
	__shared__ __align__(16) uint8_t ring[XXH_HUGE_RING_STRIDE];
	for (uint32_t i = 0; i < pl.n; i++) xxh_stream_update(st, base + pl.off[i], pl.len[i], nullptr, threadIdx.x & 31, ring);
}

__global__ void xxh32_stream_reset_kernel(XxhState *st)
{
	if (threadIdx.x < 4) st->acc[threadIdx.x] = xxh_init_acc(threadIdx.x);
	if (threadIdx.x == 0) {
		st->buf_size = 0;
		st->total_len = 0;
	}
}

// K5: size pre-pass, one thread per block: walks token / length / offset bytes only.
__global__ void __launch_bounds__(128)
size_blocks_kernel(const uint8_t *__restrict__ src, uint32_t n_blocks,
		   const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status)
{
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_blocks) return;
	const lz4b200_blk_desc d = desc[b];
	const uint8_t *s = src + d.src_off;
	const uint32_t n = d.src_len;
	uint32_t code = LZ4B200_ST_OK, err_pos = 0;
	uint64_t op = 0;
	if (d.flags & LZ4B200_BLK_STORED) {
		op = n;
	} else {
		uint32_t ip = 0;
		while (ip < n) {
			const uint32_t token = __ldg(s + ip++);
			uint32_t lit = token >> 4;
			if (lit == 15) {
				uint32_t bv;
				do {
					if (ip >= n) { code = LZ4B200_ST_LIT_EXT_OVERRUN; break; }
					bv = __ldg(s + ip++);
					lit += bv;
				} while (bv == 255);
				if (code) break;
			}
			if (lit > n - ip) {
				code = (token & 15) ? LZ4B200_ST_ENDS_AFTER_LITERALS : LZ4B200_ST_LITERAL_OVERRUN;
				break;
			}
			ip += lit;
			op += lit;
			if (ip >= n) {
				if (token & 15) code = LZ4B200_ST_ENDS_AFTER_LITERALS;
				break;
			}
			if (ip + 1 >= n) { code = LZ4B200_ST_OFFSET_TRUNCATED; break; }
			const uint32_t offset = __ldg(s + ip) | (__ldg(s + ip + 1) << 8);
			ip += 2;
			if (offset == 0) { code = LZ4B200_ST_OFFSET_ZERO; break; }
			uint32_t ml = token & 15;
			if (ml == 15) {
				uint32_t bv;
				do {
					if (ip >= n) { code = LZ4B200_ST_MATCH_EXT_OVERRUN; break; }
					bv = __ldg(s + ip++);
					ml += bv;
				} while (bv == 255);
				if (code) break;
			}
			op += ml + 4;
		}
		err_pos = op > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(op);
	}
	status[b].code = code;
	status[b].out_len = op > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(op);
	status[b].err_pos = err_pos;
	status[b].aux = 0;
	status[b].xxh32_computed = 0;
	status[b].xxh32_declared = 0;
}

// K2: stored blocks (lib/lz4ada.adb:685-695) as a wide copy -- every warp of the grid takes one 32 KiB tile of one
// stored block, so a 4 MiB block is spread over 128 warps instead of riding one warp of K1.  The block checksum, a
// serial XXH32 chain over the payload, runs beside the copy as a quad chain per block (xxh32_spans over the source).
constexpr uint32_t K2_TILE = 32768;

__global__ void __launch_bounds__(256)
copy_stored_kernel(const uint8_t *__restrict__ src, uint8_t *dst, uint32_t n_idx, const uint32_t *__restrict__ idx,
		   const lz4b200_blk_desc *__restrict__ desc, lz4b200_blk_status *status, uint32_t tiles_per_block)
{
	const int lane = threadIdx.x & 31;
	const uint64_t w = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
	const uint32_t k = static_cast<uint32_t>(w / tiles_per_block), t = static_cast<uint32_t>(w % tiles_per_block);
	if (k >= n_idx) return;
	const uint32_t b = idx[k];
	const lz4b200_blk_desc d = desc[b];
	const bool fits = d.src_len <= d.dst_cap;
	if (t == 0 && lane == 0) {
		// Check_Checksum comes first in the reference (:672-676); verify_stored_kernel overrides this status if it fails
		status[b].code = fits ? LZ4B200_ST_OK : LZ4B200_ST_OUTPUT_OVERFLOW;
		status[b].out_len = fits ? d.src_len : 0u;
		status[b].err_pos = fits ? 0u : d.src_len;
		status[b].aux = 0;
		status[b].xxh32_computed = 0;
		status[b].xxh32_declared = 0;
	}
	const uint64_t lo = static_cast<uint64_t>(t) * K2_TILE;
	if (!fits || lo >= d.src_len) return;
	const uint32_t n = d.src_len - lo < K2_TILE ? static_cast<uint32_t>(d.src_len - lo) : K2_TILE;
	warp_copy<true>(dst + d.dst_off + lo, src + d.src_off + lo, n, lane);
}

// ... and the verdict of the block checksums computed beside it (one thread per stored block with a checksum).
__global__ void __launch_bounds__(128)
verify_stored_kernel(const uint8_t *__restrict__ src, uint32_t n_idx, const uint32_t *__restrict__ idx,
		     const lz4b200_blk_desc *__restrict__ desc, const uint32_t *__restrict__ computed, lz4b200_blk_status *status)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_idx) return;
	const uint32_t b = idx[k];
	const lz4b200_blk_desc d = desc[b];
	if (!(d.flags & LZ4B200_BLK_HAS_CHECKSUM)) return;
	const uint8_t *t = src + d.src_off + d.src_len;
	const uint32_t declared = ld_u8<true>(t) | (ld_u8<true>(t + 1) << 8) | (ld_u8<true>(t + 2) << 16) | (ld_u8<true>(t + 3) << 24);
	status[b].xxh32_computed = computed[k];
	status[b].xxh32_declared = declared;
	if (computed[k] != declared) {   // lib/lz4ada.adb:702: beats everything else about the block
		status[b].code = LZ4B200_ST_BLOCK_CHECKSUM;
		status[b].out_len = 0;
		status[b].err_pos = 0;
	}
}

// The roofline denominator, measured in the same run: a plain 16-byte-per-thread grid-stride copy (read + write).
__global__ void __launch_bounds__(256)
copy_probe_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16)
{
	const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
	for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------

struct lz4b200_ctx {
	// Shared ownership: the creator holds one reference (dropped by lz4b200_destroy); every stream object,
	// decompressor, batch and the default-context slot of the LZ4Ada layer holds one more (lz4b200_retain).
	// The CUDA objects are torn down when the last reference goes, so a child outliving its creator's
	// lz4b200_destroy keeps working instead of touching freed memory.
	std::atomic<int> refs{1};
	// host-layer scratch kept per context (batch.cpp's pool): freed with the context
	std::mutex attach_mutex;
	void *attachment = nullptr;
	void (*attachment_free)(lz4b200_ctx *, void *) = nullptr;
	void *attachment2 = nullptr;   // a second host-layer object (the streaming engine's buffer pool)
	void (*attachment2_free)(lz4b200_ctx *, void *) = nullptr;
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;      // the lane in use (lz4b200_use_lane)
	cudaStream_t lanes[4] = {nullptr, nullptr, nullptr, nullptr};   // lane 0 = primary stream
	bool own_stream = false;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	uint64_t launches = 0;
	uint32_t *d_counter = nullptr;          // v5: block queue heads, one per lane (stream) of the context
	unsigned long long *d_prof = nullptr;   // LZ4B200_PROF=1: v3 phase counters (printed by lz4b200_destroy)
	int blocks_per_warp = 0;   // K1 tuning: 0 = auto (v5 for big batches, else v4), 50 = v5, 40..48 = v4, 64 = v3,
	                           // 1..16 = v2 with G blocks per warp, -1 = v1 kernel
	char err[256] = "";
};

static int fail(lz4b200_ctx *ctx, cudaError_t e, const char *what)
{
	if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, cudaGetErrorString(e));
	return LZ4B200_ERR_CUDA;
}

#define CK(call)                                                                        \
	do {                                                                            \
		cudaError_t e_ = (call);                                                \
		if (e_ != cudaSuccess) return fail(ctx, e_, #call);                     \
	} while (0)

// Host-layer attachment (not part of the C-ABI: C++ linkage, see host/common.hpp).
namespace lz4ada {
void *ctx_attachment(lz4b200_ctx *ctx, void *(*make)(lz4b200_ctx *), void (*free_fn)(lz4b200_ctx *, void *))
{
	if (!ctx) return nullptr;
	std::lock_guard<std::mutex> lock(ctx->attach_mutex);
	if (!ctx->attachment && make) {
		ctx->attachment = make(ctx);
		ctx->attachment_free = free_fn;
	}
	return ctx->attachment;
}
void *ctx_attachment2(lz4b200_ctx *ctx, void *(*make)(lz4b200_ctx *), void (*free_fn)(lz4b200_ctx *, void *))
{
	if (!ctx) return nullptr;
	std::lock_guard<std::mutex> lock(ctx->attach_mutex);
	if (!ctx->attachment2 && make) {
		ctx->attachment2 = make(ctx);
		ctx->attachment2_free = free_fn;
	}
	return ctx->attachment2;
}
}  // namespace lz4ada

extern "C" {

int lz4b200_create(int device, void *stream, lz4b200_ctx **out)
{
	if (!out) return LZ4B200_ERR_ARG;
	*out = nullptr;
	lz4b200_ctx *ctx = new (std::nothrow) lz4b200_ctx();
	if (!ctx) return LZ4B200_ERR_NOMEM;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || device < 0 || device >= ndev) {
		// no CUDA device: there is deliberately no CPU fallback
		delete ctx;
		return LZ4B200_ERR_CUDA;
	}
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return LZ4B200_ERR_CUDA; }
	cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
	if (stream) {
		ctx->stream = static_cast<cudaStream_t>(stream);
	} else {
		if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
			delete ctx;
			return LZ4B200_ERR_CUDA;
		}
		ctx->own_stream = true;
	}
	ctx->lanes[0] = ctx->stream;
	{
		// once per context, not per launch: the K3 kernels need > 48 KiB of dynamic shared memory
		const int smem = 4 * 8 * XXH_RING_STRIDE;
		cudaFuncSetAttribute(xxh32_frames_kernel<XXH_GROUP_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		cudaFuncSetAttribute(xxh32_frames_kernel<XXH_BIG_GROUP_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
				     int(8 * XXH_BIG_RING_STRIDE));
		cudaFuncSetAttribute(xxh32_frames_kernel<XXH_HUGE_GROUP_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
				     int(8 * XXH_HUGE_RING_STRIDE));
		cudaFuncSetAttribute(xxh32_frames_kernel<XXH_MID_GROUP_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
				     int(8 * XXH_MID_RING_STRIDE));
		cudaFuncSetAttribute(xxh32_spans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		cudaFuncSetAttribute(decode_blocks_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(v3::SMEM_BYTES));
		cudaFuncSetAttribute(decode_blocks_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
				     int(v5::WARPS * sizeof(v5::WarpMem)));
		cudaFuncSetAttribute(decode_blocks_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
				     int(v4::WARPS * sizeof(v4::WarpMem)));
		cudaFuncSetAttribute(decode_chain_k6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(K6_SMEM));
		cudaFuncSetAttribute(decode_chain_k7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(K7_SMEM));
		cudaFuncSetAttribute(stream_block_k7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(K7_SMEM));
		cudaFuncSetAttribute(decode_blocks_v6_kernel<64, V6A_K, V6A_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
				     int(v6::Layout<64, V6A_K>::smem_bytes(V6A_WARPS)));
		cudaFuncSetAttribute(decode_blocks_v6_kernel<128, V6B_K, V6B_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
				     int(v6::Layout<128, V6B_K>::smem_bytes(V6B_WARPS)));
	}
	cudaEventCreate(&ctx->ev0);
	cudaEventCreate(&ctx->ev1);
	if (cudaMalloc(&ctx->d_counter, 4 * 64) != cudaSuccess) ctx->d_counter = nullptr;
	if (const char *e = getenv("LZ4B200_PROF")) {
		if (e[0] == '1' && cudaMalloc(&ctx->d_prof, sizeof(unsigned long long) * v3::PROF_N) == cudaSuccess)
			cudaMemset(ctx->d_prof, 0, sizeof(unsigned long long) * v3::PROF_N);
	}
	*out = ctx;
	return LZ4B200_OK;
}

int lz4b200_retain(lz4b200_ctx *ctx)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	ctx->refs.fetch_add(1, std::memory_order_relaxed);
	return LZ4B200_OK;
}

int lz4b200_destroy(lz4b200_ctx *ctx)
{
	if (!ctx) return LZ4B200_OK;
	if (ctx->refs.fetch_sub(1, std::memory_order_acq_rel) > 1) return LZ4B200_OK;   // children still hold it
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->lanes[0]);
	if (ctx->attachment && ctx->attachment_free) ctx->attachment_free(ctx, ctx->attachment);
	ctx->attachment = nullptr;
	if (ctx->attachment2 && ctx->attachment2_free) ctx->attachment2_free(ctx, ctx->attachment2);
	ctx->attachment2 = nullptr;
	if (ctx->d_prof) {
		static const char *names[v3::PROF_N] = {"load", "parse", "scan", "emit", "match", "flush", "exact", "blocks", "windows",
							"iters", "fallback", "batches", "rounds", "early", "gatewait", "coop"};
		unsigned long long h[v3::PROF_N] = {};
		cudaDeviceSynchronize();
		cudaMemcpy(h, ctx->d_prof, sizeof h, cudaMemcpyDeviceToHost);
		fprintf(stderr, "[lz4b200 v3 prof]");
		for (int i = 0; i < v3::PROF_N; i++) fprintf(stderr, " %s=%llu", names[i], h[i]);
		fprintf(stderr, "\n");
		cudaFree(ctx->d_prof);
	}
	if (ctx->d_counter) cudaFree(ctx->d_counter);
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	for (int i = 1; i < 4; i++)
		if (ctx->lanes[i]) {
			cudaStreamSynchronize(ctx->lanes[i]);
			cudaStreamDestroy(ctx->lanes[i]);
		}
	if (ctx->own_stream) cudaStreamDestroy(ctx->lanes[0]);
	delete ctx;
	return LZ4B200_OK;
}

const char *lz4b200_last_error(const lz4b200_ctx *ctx) { return ctx ? ctx->err : "no context"; }
int lz4b200_sm_count(const lz4b200_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t lz4b200_launch_count(const lz4b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int lz4b200_use_lane(lz4b200_ctx *ctx, int lane)
{
	if (!ctx || lane < 0 || lane >= 4) return LZ4B200_ERR_ARG;
	if (!ctx->lanes[lane]) {
		CK(cudaSetDevice(ctx->device));
		CK(cudaStreamCreateWithFlags(&ctx->lanes[lane], cudaStreamNonBlocking));
	}
	ctx->stream = ctx->lanes[lane];
	return LZ4B200_OK;
}

int lz4b200_sync_all(lz4b200_ctx *ctx)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	for (int i = 0; i < 4; i++)
		if (ctx->lanes[i]) CK(cudaStreamSynchronize(ctx->lanes[i]));
	return LZ4B200_OK;
}

int lz4b200_set_tuning(lz4b200_ctx *ctx, int blocks_per_warp)
{
	if (!ctx || (blocks_per_warp != -1 && blocks_per_warp != 0 && blocks_per_warp != 1 && blocks_per_warp != 2 &&
		     blocks_per_warp != 4 && blocks_per_warp != 8 && blocks_per_warp != 16 && blocks_per_warp != 64 && blocks_per_warp != 40 && blocks_per_warp != 41 &&
		     blocks_per_warp != 42 && blocks_per_warp != 44 && blocks_per_warp != 48 && blocks_per_warp != 50 && blocks_per_warp != 60 &&
		     blocks_per_warp != 61))
		return LZ4B200_ERR_ARG;
	ctx->blocks_per_warp = blocks_per_warp;
	return LZ4B200_OK;
}

int lz4b200_get_tuning(const lz4b200_ctx *ctx) { return ctx ? ctx->blocks_per_warp : 0; }

int lz4b200_k1_fallbacks(lz4b200_ctx *ctx, uint32_t *to_exact, uint32_t *by_safety_net)
{
	if (!ctx || !ctx->d_counter) return LZ4B200_ERR_ARG;
	uint32_t h[6] = {0, 0, 0, 0, 0, 0};
	int li = 0;
	for (int i = 0; i < 4; i++)
		if (ctx->lanes[i] == ctx->stream) li = i;
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaMemcpy(h, ctx->d_counter + 16 * li, sizeof h, cudaMemcpyDeviceToHost));
	if (to_exact) *to_exact = h[1];
	if (by_safety_net) *by_safety_net = h[2];
	if (getenv("LZ4B200_V6_DEBUG") && h[1]) fprintf(stderr, "[lz4b200 v6] %u blocks to the exact routine; last: block %u, reason bits 0x%x, input position %u\n", h[1], h[3], h[4], h[5]);
	return LZ4B200_OK;
}

int lz4b200_chain_stats(lz4b200_ctx *ctx, uint32_t *finished, uint32_t *given_up)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	uint32_t h[8] = {};
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaMemcpyFromSymbol(h, k7::g_stats, sizeof h));
	if (finished) *finished = h[0];
	if (given_up) *given_up = h[1];
	memset(h, 0, sizeof h);
	CK(cudaMemcpyToSymbol(k7::g_stats, h, sizeof h));
	return LZ4B200_OK;
}

// the auto rule of lz4b200_decode_blocks, in one place
static int k1_generation(const lz4b200_ctx *ctx, uint32_t n_blocks)
{
	int g = ctx->blocks_per_warp;
	if (g == 0) {
		// v6 (a lane per block) runs a batch in the time one block takes when the batch fits the resident lanes
		// (148 SMs x 14 warps x 32), and that time does not shrink with the batch: ~5 ms for 64 KiB text blocks.
		// v4 (a warp per block) costs ~0.37 us per such block, so it is the faster shape below ~20 000 blocks
		// (measured: 16 384 blocks 6.0 ms v4 / 7.2 ms v6, 32 768 blocks 11.6 / 7.8 ms).
		g = n_blocks >= 20000u ? 60 : 40;
	}
	return g;
}

const char *lz4b200_k1_kernel_name(const lz4b200_ctx *ctx, uint32_t n_blocks)
{
	if (!ctx) return "";
	const int g = k1_generation(ctx, n_blocks);
	return g == 60 || g == 61 ? "decode_blocks_v6_kernel" : g == 50 ? "decode_blocks_v5_kernel" : g >= 40 && g <= 48 ? "decode_blocks_v4_kernel" : g == 64 ? "decode_blocks_v3_kernel"
	       : g < 0 ? "decode_blocks_kernel" : "decode_blocks_v2_kernel";
}

int lz4b200_alloc(lz4b200_ctx *ctx, size_t bytes, void **dev_ptr)
{
	if (!ctx || !dev_ptr) return LZ4B200_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	CK(cudaMalloc(dev_ptr, bytes + 64));   // slack: aligned 16-byte loads may touch the granule past the end
	return LZ4B200_OK;
}

int lz4b200_free(lz4b200_ctx *ctx, void *dev_ptr)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	CK(cudaFree(dev_ptr));
	return LZ4B200_OK;
}

int lz4b200_alloc_host(lz4b200_ctx *ctx, size_t bytes, void **host_ptr)
{
	if (!ctx || !host_ptr) return LZ4B200_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	CK(cudaMallocHost(host_ptr, bytes ? bytes : 1));
	return LZ4B200_OK;
}

int lz4b200_free_host(lz4b200_ctx *ctx, void *host_ptr)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	CK(cudaFreeHost(host_ptr));
	return LZ4B200_OK;
}

int lz4b200_h2d(lz4b200_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (bytes == 0) return LZ4B200_OK;
	CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return LZ4B200_OK;
}

int lz4b200_d2h(lz4b200_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (bytes == 0) return LZ4B200_OK;
	CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	return LZ4B200_OK;
}

int lz4b200_memset(lz4b200_ctx *ctx, void *dst_dev, int value, size_t bytes)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (bytes == 0) return LZ4B200_OK;
	CK(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
	return LZ4B200_OK;
}

int lz4b200_sync(lz4b200_ctx *ctx)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	CK(cudaStreamSynchronize(ctx->stream));
	return LZ4B200_OK;
}

int lz4b200_timer_start(lz4b200_ctx *ctx)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	CK(cudaEventRecord(ctx->ev0, ctx->stream));
	return LZ4B200_OK;
}

int lz4b200_timer_stop(lz4b200_ctx *ctx, float *elapsed_ms)
{
	if (!ctx || !elapsed_ms) return LZ4B200_ERR_ARG;
	CK(cudaEventRecord(ctx->ev1, ctx->stream));
	CK(cudaEventSynchronize(ctx->ev1));
	CK(cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
	return LZ4B200_OK;
}

int lz4b200_copy_probe(lz4b200_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes, int reps, float *best_ms)
{
	if (!ctx || !dst_dev || !src_dev || !best_ms || reps < 1 || bytes < 16) return LZ4B200_ERR_ARG;
	const size_t n16 = bytes / 16;
	const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
	float best = 0;
	for (int r = 0; r < reps + 1; r++) {   // (the first pass warms up)
		CK(cudaEventRecord(ctx->ev0, ctx->stream));
		copy_probe_kernel<<<sms * 16, 256, 0, ctx->stream>>>(static_cast<const uint4 *>(src_dev), static_cast<uint4 *>(dst_dev), n16);
		ctx->launches++;
		CK(cudaGetLastError());
		CK(cudaEventRecord(ctx->ev1, ctx->stream));
		CK(cudaEventSynchronize(ctx->ev1));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		if (r > 0 && (best == 0 || ms < best)) best = ms;
	}
	*best_ms = best;
	return LZ4B200_OK;
}

int lz4b200_device_of(const lz4b200_ctx *ctx) { return ctx ? ctx->device : -1; }

int lz4b200_event_create(lz4b200_ctx *ctx, void **event)
{
	if (!ctx || !event) return LZ4B200_ERR_ARG;
	cudaEvent_t e;
	CK(cudaEventCreate(&e));
	*event = e;
	return LZ4B200_OK;
}

int lz4b200_event_destroy(lz4b200_ctx *ctx, void *event)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (event) CK(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
	return LZ4B200_OK;
}

int lz4b200_event_record(lz4b200_ctx *ctx, void *event)
{
	if (!ctx || !event) return LZ4B200_ERR_ARG;
	CK(cudaEventRecord(static_cast<cudaEvent_t>(event), ctx->stream));
	return LZ4B200_OK;
}

int lz4b200_event_sync(lz4b200_ctx *ctx, void *event)
{
	if (!ctx || !event) return LZ4B200_ERR_ARG;
	CK(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
	return LZ4B200_OK;
}

int lz4b200_event_elapsed(lz4b200_ctx *ctx, void *start, void *stop, float *elapsed_ms)
{
	if (!ctx || !start || !stop || !elapsed_ms) return LZ4B200_ERR_ARG;
	CK(cudaEventElapsedTime(elapsed_ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
	return LZ4B200_OK;
}

int lz4b200_decode_blocks(lz4b200_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t n_blocks,
			  const lz4b200_blk_desc *desc, lz4b200_blk_status *status)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (n_blocks == 0) return LZ4B200_OK;
	int g = k1_generation(ctx, n_blocks);
	if (g == 64) {
		// v3 (kept selectable: measured slower than v2): up to eight blocks per CTA (one hash chain per quad), but at least ~8 waves of CTAs
		const uint32_t slots = static_cast<uint32_t>(ctx->sm_count > 0 ? ctx->sm_count : 148) * 2u;
		uint32_t per = n_blocks / (8u * slots);
		per = per < 1 ? 1 : per > v3::MAX_NB ? v3::MAX_NB : per;
		const uint32_t grid = (n_blocks + per - 1) / per;
		decode_blocks_v3_kernel<<<grid, v3::CTA_THREADS, v3::SMEM_BYTES, ctx->stream>>>(src, dst, n_blocks, desc, status, per, ctx->d_prof);
		ctx->launches++;
		CK(cudaGetLastError());
		return LZ4B200_OK;
	}
	if (g == 60 || g == 61) {
		// v6: persistent warps whose lanes pull blocks from a counter (one counter per stream lane of the context)
		if (!ctx->d_counter) return LZ4B200_ERR_NOMEM;
		int li = 0;
		for (int i = 0; i < 4; i++)
			if (ctx->lanes[i] == ctx->stream) li = i;
		uint32_t *counter = ctx->d_counter + 16 * li;
		CK(cudaMemsetAsync(counter, 0, 24, ctx->stream));   // [0] block queue head, [1] blocks handed to the exact routine, [2] of those by the safety net, [3..5] the last of them: block, reason bits, input position
		const uint32_t sms = static_cast<uint32_t>(ctx->sm_count > 0 ? ctx->sm_count : 148);
		// One CTA per SM.  A block is one lane's work from start to end, so a batch that fits the resident lanes runs
		// as long as one block does: give each SM just the warps that hold the batch (the fewer share an SM, the
		// shorter their trips); more blocks than 148 x WARPS x 32 lanes go round in turns.
		const uint32_t warps = (n_blocks + 31) / 32;
		const uint32_t max_w = g == 60 ? V6A_WARPS : V6B_WARPS;
		uint32_t wpc = (warps + sms - 1) / sms;
		wpc = wpc < 2 ? 2 : wpc > max_w ? max_w : wpc;
		if (const char *e = getenv("LZ4B200_V6_WARPS")) {
			const uint32_t w = static_cast<uint32_t>(atoi(e));
			if (w >= 1 && w <= max_w) wpc = w;
		}
		uint32_t grid = (warps + wpc - 1) / wpc;
		if (grid > sms) grid = sms;
		if (g == 60)
			decode_blocks_v6_kernel<64, V6A_K, V6A_WARPS><<<grid, wpc * 32, v6::Layout<64, V6A_K>::smem_bytes(wpc), ctx->stream>>>(
				src, dst, n_blocks, desc, status, counter);
		else
			decode_blocks_v6_kernel<128, V6B_K, V6B_WARPS><<<grid, wpc * 32, v6::Layout<128, V6B_K>::smem_bytes(wpc), ctx->stream>>>(
				src, dst, n_blocks, desc, status, counter);
		ctx->launches++;
		CK(cudaGetLastError());
		return LZ4B200_OK;
	}
	if (g == 50) {
		// v5: lanes pull blocks from a counter; one counter per stream lane of the context
		if (!ctx->d_counter) return LZ4B200_ERR_NOMEM;
		int li = 0;
		for (int i = 0; i < 4; i++)
			if (ctx->lanes[i] == ctx->stream) li = i;
		uint32_t *counter = ctx->d_counter + 16 * li;
		CK(cudaMemsetAsync(counter, 0, 4, ctx->stream));
		const uint32_t sms = static_cast<uint32_t>(ctx->sm_count > 0 ? ctx->sm_count : 148);
		uint32_t warps = (n_blocks + 31) / 32;
		if (warps > sms * 16) warps = sms * 16;
		const uint32_t grid = (warps + v5::WARPS - 1) / v5::WARPS;
		decode_blocks_v5_kernel<<<grid, v5::WARPS * 32, v5::WARPS * sizeof(v5::WarpMem), ctx->stream>>>(src, dst, n_blocks, desc,
														status, counter);
		ctx->launches++;
		CK(cudaGetLastError());
		return LZ4B200_OK;
	}
	if (g >= 40 && g <= 48) {
		// v4: G blocks per warp (40 = choose: enough warps for ~4 waves of 20 warps per SM)
		uint32_t G = static_cast<uint32_t>(g - 40);
		if (G == 0) {
			const uint32_t per = static_cast<uint32_t>(ctx->sm_count > 0 ? ctx->sm_count : 148) * 20u * 4u;
			G = n_blocks >= 8 * per ? 8 : n_blocks >= 4 * per ? 4 : n_blocks >= 2 * per ? 2 : 1;
		}
		const uint32_t warps = (n_blocks + G - 1) / G;
		// A CTA only leaves the SM when all its warps are done, and blocks differ wildly in how long they take (a
		// 4 MiB text block 40 ms, an RLE or stored one microseconds): with four warps per CTA the 16 GiB mixed corpus
		// (4096 blocks) ran in two waves, every CTA held hostage by its one text block.  Few warps => one per CTA;
		// many (small blocks, short-lived) => four per CTA, which launches faster.  LZ4B200_V4_WARPS=1|4 forces it.
		static const int wpc_env = [] {
			const char *e = getenv("LZ4B200_V4_WARPS");
			return e ? (e[0] == '4' ? int(v4::WARPS) : 1) : 0;
		}();
		const int wpc = wpc_env ? wpc_env : (warps <= 8192u ? 1 : int(v4::WARPS));
		const uint32_t grid = (warps + wpc - 1) / wpc;
		decode_blocks_v4_kernel<<<grid, wpc * 32, wpc * sizeof(v4::WarpMem), ctx->stream>>>(src, dst, n_blocks, desc, status, G);
		ctx->launches++;
		CK(cudaGetLastError());
		return LZ4B200_OK;
	}
	if (g == 0) {
		// keep at least ~16 warps per SM busy; more blocks per warp = cheaper token-chain walking
		const uint32_t per = static_cast<uint32_t>(ctx->sm_count > 0 ? ctx->sm_count : 148) * 16u;
		g = n_blocks >= 16 * per + per ? 16 : n_blocks >= 8 * per ? 8 : n_blocks >= 4 * per ? 4 : n_blocks >= 2 * per ? 2 : 1;
	}
	if (g < 0) {
		const uint32_t grid = (n_blocks + K1_WARPS - 1) / K1_WARPS;
		decode_blocks_kernel<<<grid, K1_WARPS * 32, 0, ctx->stream>>>(src, dst, n_blocks, desc, status);
	} else {
		const uint32_t warps = (n_blocks + g - 1) / g;
		const uint32_t grid = (warps + K1_WARPS - 1) / K1_WARPS;
		switch (g) {
		case 16: decode_blocks_v2_kernel<16><<<grid, K1_WARPS * 32, 0, ctx->stream>>>(src, dst, n_blocks, desc, status); break;
		case 8: decode_blocks_v2_kernel<8><<<grid, K1_WARPS * 32, 0, ctx->stream>>>(src, dst, n_blocks, desc, status); break;
		case 4: decode_blocks_v2_kernel<4><<<grid, K1_WARPS * 32, 0, ctx->stream>>>(src, dst, n_blocks, desc, status); break;
		case 2: decode_blocks_v2_kernel<2><<<grid, K1_WARPS * 32, 0, ctx->stream>>>(src, dst, n_blocks, desc, status); break;
		default: decode_blocks_v2_kernel<1><<<grid, K1_WARPS * 32, 0, ctx->stream>>>(src, dst, n_blocks, desc, status); break;
		}
	}
	ctx->launches++;
	CK(cudaGetLastError());
	return LZ4B200_OK;
}

int lz4b200_copy_stored(lz4b200_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t n_idx, const uint32_t *idx,
			uint32_t max_len, const lz4b200_blk_desc *desc, lz4b200_blk_status *status,
			const lz4b200_hash_span *spans, uint32_t *scratch)
{
	if (!ctx || (n_idx && !idx)) return LZ4B200_ERR_ARG;
	if (n_idx == 0) return LZ4B200_OK;
	const uint32_t tiles = max_len ? (max_len + K2_TILE - 1) / K2_TILE : 1;
	const uint64_t warps = static_cast<uint64_t>(n_idx) * tiles;
	const uint64_t grid = (warps + 7) / 8;
	if (grid > 0x7fffffffull) return LZ4B200_ERR_ARG;
	copy_stored_kernel<<<static_cast<uint32_t>(grid), 256, 0, ctx->stream>>>(src, dst, n_idx, idx, desc, status, tiles);
	ctx->launches++;
	CK(cudaGetLastError());
	if (spans && scratch) {
		// block checksums of the stored payloads: one quad chain per block, then the comparison
		const int rc = lz4b200_xxh32_spans(ctx, src, n_idx, spans, scratch);
		if (rc != LZ4B200_OK) return rc;
		verify_stored_kernel<<<(n_idx + 127) / 128, 128, 0, ctx->stream>>>(src, n_idx, idx, desc, scratch, status);
		ctx->launches++;
		CK(cudaGetLastError());
	}
	return LZ4B200_OK;
}

int lz4b200_decode_linked(lz4b200_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t n_chains,
			  const lz4b200_chain *chains, const lz4b200_blk_desc *desc,
			  lz4b200_blk_status *status)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (n_chains == 0) return LZ4B200_OK;
	// The default is K7 (one CTA per chain: speculative parse + pointer jumping, kernels_k7.cuh).  For A/B comparisons and
	// tests LZ4B200_CHAIN_KERNEL selects the earlier generations: pipe = K4 (one parser warp, seven gated copier warps),
	// k6 = K6 (pointer-doubling parser + round-based copier on a shared-memory window), warp = one warp per chain
	static const int which = [] {
		const char *e = getenv("LZ4B200_CHAIN_KERNEL");
		return e && e[0] == 'w' ? 1 : e && e[0] == 'k' && e[1] == '6' ? 0 : e && e[0] == 'p' ? 2 : 3;
	}();
	if (which == 3) {
		chain_block_hashes_kernel<<<n_chains, 128, 0, ctx->stream>>>(src, n_chains, chains, desc, status);
		decode_chain_k7_kernel<<<n_chains, k7::T, K7_SMEM, ctx->stream>>>(src, dst, n_chains, chains, desc, status);
		ctx->launches += 2;
		CK(cudaGetLastError());
		static const bool dbg7 = getenv("LZ4B200_K7_DEBUG") != nullptr;
		if (dbg7) {
			uint32_t h[8] = {};
			CK(cudaStreamSynchronize(ctx->stream));
			CK(cudaMemcpyFromSymbol(h, k7::g_stats, sizeof h));
			fprintf(stderr, "[lz4b200 k7] blocks finished %u, given up %u (last: reason %u, block %u, input position %u)\n", h[0], h[1], h[2], h[3], h[4]);
			memset(h, 0, sizeof h);
			CK(cudaMemcpyToSymbol(k7::g_stats, h, sizeof h));
			unsigned long long q[16] = {};
			CK(cudaMemcpyFromSymbol(q, k7::g_prof, sizeof q));
			fprintf(stderr, "[lz4b200 k7] steps %llu, parse iterations %llu, resolve calls %llu rounds %llu; kilocycles (thread 0, summed over CTAs): stage %llu parse %llu place %llu windows %llu (init %llu resolve %llu flush %llu); parse = table %llu + first walk %llu + neighbours %llu + path %llu\n",
				q[0], q[1], q[2], q[3], q[4] >> 10, q[5] >> 10, q[6] >> 10, q[7] >> 10, q[10] >> 10, q[8] >> 10, q[9] >> 10, q[11] >> 10, q[12] >> 10, q[13] >> 10, q[14] >> 10);
			memset(q, 0, sizeof q);
			CK(cudaMemcpyToSymbol(k7::g_prof, q, sizeof q));
		}
		return LZ4B200_OK;
	}
	if (which == 1) {
		const uint32_t grid = (n_chains + K1_WARPS - 1) / K1_WARPS;
		decode_linked_kernel<<<grid, K1_WARPS * 32, 0, ctx->stream>>>(src, dst, n_chains, chains, desc, status);
	} else if (which == 0) {
		static const uint32_t dbg = [] {
			const char *e = getenv("LZ4B200_K6_DBG");   // 1: the copier skips its copies (timing of the parser alone; output is wrong)
			return e ? static_cast<uint32_t>(atoi(e)) : 0u;
		}();
		decode_chain_k6_kernel<<<n_chains, 64, K6_SMEM, ctx->stream>>>(src, dst, n_chains, chains, desc, status, dbg);
	} else {
		decode_chain_pipe_kernel<<<n_chains, PIPE_WARPS * 32, 0, ctx->stream>>>(src, dst, n_chains, chains, desc, status);
	}
	ctx->launches++;
	CK(cudaGetLastError());
	return LZ4B200_OK;
}

int lz4b200_xxh32_frames(lz4b200_ctx *ctx, const uint8_t *dst, uint32_t n_frames,
			 const lz4b200_frame_blocks *frames, const lz4b200_blk_desc *desc,
			 const lz4b200_blk_status *status, uint32_t *digest, uint32_t *valid)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (n_frames == 0) return LZ4B200_OK;
	const uint32_t warps = (n_frames + 7) / 8;
	// A chain needs 16 bytes every ~14 cycles and a quad can only keep its ring in flight.  With thousands of
	// frames the sum is plenty (4096 frames x 2 KiB: 2.9 TB/s); with few, long ones (1024 frames of 4 MiB: one
	// warp per SM) the 2 KiB ring is the limit (0.75 TB/s), so those get 8 or 16 KiB rings, one warp per CTA.
	const uint32_t sms = static_cast<uint32_t>(ctx->sm_count > 0 ? ctx->sm_count : 148);
	if (warps <= sms) {
		xxh32_frames_kernel<XXH_HUGE_GROUP_BYTES><<<warps, 32, 8 * XXH_HUGE_RING_STRIDE, ctx->stream>>>(dst, n_frames, frames, desc,
														status, digest, valid);
	} else if (warps <= 3 * sms) {
		xxh32_frames_kernel<XXH_BIG_GROUP_BYTES><<<warps, 32, 8 * XXH_BIG_RING_STRIDE, ctx->stream>>>(dst, n_frames, frames, desc,
													      status, digest, valid);
	} else if (warps <= 6 * sms) {
		xxh32_frames_kernel<XXH_MID_GROUP_BYTES><<<warps, 32, 8 * XXH_MID_RING_STRIDE, ctx->stream>>>(dst, n_frames, frames, desc,
													      status, digest, valid);
	} else {
		const size_t smem = 4 * 8 * XXH_RING_STRIDE;
		xxh32_frames_kernel<XXH_GROUP_BYTES><<<(warps + 3) / 4, 128, smem, ctx->stream>>>(dst, n_frames, frames, desc, status,
												   digest, valid);
	}
	ctx->launches++;
	CK(cudaGetLastError());
	return LZ4B200_OK;
}

int lz4b200_xxh32_spans(lz4b200_ctx *ctx, const uint8_t *data, uint32_t n,
			const lz4b200_hash_span *spans, uint32_t *out)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (n == 0) return LZ4B200_OK;
	const uint32_t warps = (n + 7) / 8;
	const uint32_t grid = (warps + 3) / 4;
	const size_t smem = 4 * 8 * XXH_RING_STRIDE;
	xxh32_spans_kernel<<<grid, 128, smem, ctx->stream>>>(data, n, spans, out);
	ctx->launches++;
	CK(cudaGetLastError());
	return LZ4B200_OK;
}

int lz4b200_size_blocks(lz4b200_ctx *ctx, const uint8_t *src, uint32_t n_blocks,
			const lz4b200_blk_desc *desc, lz4b200_blk_status *status)
{
	if (!ctx) return LZ4B200_ERR_ARG;
	if (n_blocks == 0) return LZ4B200_OK;
	size_blocks_kernel<<<(n_blocks + 127) / 128, 128, 0, ctx->stream>>>(src, n_blocks, desc, status);
	ctx->launches++;
	CK(cudaGetLastError());
	return LZ4B200_OK;
}

// ------------------------------------------------------------------------------------------
// Single-block streaming path under Decompressor.Update
// ------------------------------------------------------------------------------------------

struct lz4b200_stream {
	lz4b200_ctx *ctx = nullptr;
	uint32_t max_block = 0;        // largest payload / output of one call
	uint8_t *d_src = nullptr;      // payload (+4 checksum bytes) of the current block
	uint8_t *d_win = nullptr;      // [history | blocks ...] flat window
	size_t win_size = 0;
	size_t cursor = 0;             // where the next block's output goes inside d_win
	uint64_t frame_pos = 0;        // bytes of the current frame produced so far
	uint8_t *d_meta = nullptr;     // blk_desc | blk_status | XxhState | digest
	uint8_t *h_meta = nullptr;     // pinned mirror
};

constexpr size_t HISTORY = 65536;
constexpr size_t META_DESC = 0, META_STATUS = 64, META_XXH = 128, META_DIGEST = 192, META_BYTES = 256;

int lz4b200_stream_create(lz4b200_ctx *ctx, uint32_t max_block, lz4b200_stream **out)
{
	if (!ctx || !out) return LZ4B200_ERR_ARG;
	*out = nullptr;
	lz4b200_stream *s = new (std::nothrow) lz4b200_stream();
	if (!s) return LZ4B200_ERR_NOMEM;
	s->ctx = ctx;
	lz4b200_retain(ctx);
	s->max_block = max_block;
	s->win_size = HISTORY + 2 * static_cast<size_t>(max_block) + 256;
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_src, static_cast<size_t>(max_block) + 64);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_win, s->win_size + 64);
	if (e == cudaSuccess) e = cudaMalloc(&s->d_meta, META_BYTES);
	if (e == cudaSuccess) e = cudaMallocHost(&s->h_meta, META_BYTES);
	if (e == cudaSuccess) {
		s->cursor = HISTORY;
		xxh32_stream_reset_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<XxhState *>(s->d_meta + META_XXH));
		ctx->launches++;
		e = cudaGetLastError();
	}
	if (e != cudaSuccess) {
		const int rc = fail(ctx, e, "lz4b200_stream_create");
		lz4b200_stream_destroy(s);
		return rc;
	}
	*out = s;
	return LZ4B200_OK;
}

int lz4b200_stream_destroy(lz4b200_stream *s)
{
	if (!s) return LZ4B200_OK;
	cudaSetDevice(s->ctx->device);
	cudaStreamSynchronize(s->ctx->stream);
	cudaFree(s->d_src);
	cudaFree(s->d_win);
	cudaFree(s->d_meta);
	cudaFreeHost(s->h_meta);
	lz4b200_destroy(s->ctx);   // drops the stream's reference
	delete s;
	return LZ4B200_OK;
}

int lz4b200_stream_reset(lz4b200_stream *s)
{
	if (!s) return LZ4B200_ERR_ARG;
	lz4b200_ctx *ctx = s->ctx;
	s->cursor = HISTORY;
	s->frame_pos = 0;
	xxh32_stream_reset_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<XxhState *>(s->d_meta + META_XXH));
	ctx->launches++;
	CK(cudaGetLastError());
	return LZ4B200_OK;
}

int lz4b200_stream_block2(lz4b200_stream *s, const uint8_t *host_src, uint32_t src_len, uint32_t flags,
			  int hash_content, uint8_t *host_dst, uint32_t dst_cap, uint32_t expect_out,
			  lz4b200_blk_status *status)
{
	if (!s || !status) return LZ4B200_ERR_ARG;
	lz4b200_ctx *ctx = s->ctx;
	const uint32_t trailer = (flags & LZ4B200_BLK_HAS_CHECKSUM) ? 4u : 0u;
	if (static_cast<uint64_t>(src_len) + trailer > static_cast<uint64_t>(s->max_block) + 8 ||
	    dst_cap > s->max_block)
		return LZ4B200_ERR_ARG;
	// keep the last 64 KiB in front of the cursor; compact when the window is exhausted
	if (s->cursor + dst_cap > s->win_size) {
		CK(cudaMemcpyAsync(s->d_win, s->d_win + s->cursor - HISTORY, HISTORY, cudaMemcpyDeviceToDevice,
				   ctx->stream));
		s->cursor = HISTORY;
	}
	CK(cudaMemcpyAsync(s->d_src, host_src, static_cast<size_t>(src_len) + trailer, cudaMemcpyHostToDevice,
			   ctx->stream));
	lz4b200_blk_desc *hd = reinterpret_cast<lz4b200_blk_desc *>(s->h_meta + META_DESC);
	hd->src_off = 0;
	hd->dst_off = s->cursor;
	hd->src_len = src_len;
	hd->dst_cap = dst_cap;
	hd->flags = flags;
	hd->hist_avail = s->frame_pos > 0xfffffffeull ? 0xffffffffu : static_cast<uint32_t>(s->frame_pos);
	CK(cudaMemcpyAsync(s->d_meta + META_DESC, hd, sizeof *hd, cudaMemcpyHostToDevice, ctx->stream));
	// the history in front of the cursor is final: matches may read backwards across the
	// block boundary (ALLOW_HIST)
	lz4b200_blk_status *d_st = reinterpret_cast<lz4b200_blk_status *>(s->d_meta + META_STATUS);
	// a block of some size gets a CTA (the chain kernel's parallel parse and pointer jumping), a small one a warp
	static const bool big_on = [] {
		const char *e = getenv("LZ4B200_STREAM_K7");
		return !(e && e[0] == '0');
	}();
	if (big_on && src_len >= 16384u && !(flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_HASH_ONLY)))
		stream_block_k7_kernel<<<1, k7::T, K7_SMEM, ctx->stream>>>(
			s->d_src, s->d_win, reinterpret_cast<const lz4b200_blk_desc *>(s->d_meta + META_DESC), d_st);
	else
		stream_block_kernel<<<1, 32, 0, ctx->stream>>>(
			s->d_src, s->d_win, reinterpret_cast<const lz4b200_blk_desc *>(s->d_meta + META_DESC), d_st);
	ctx->launches++;
	CK(cudaGetLastError());
	// One synchronisation per block when the caller can say how much a well-formed block produces at most (the
	// frame's block maximum, up to 256 KiB): the content hash reads the block's status on the device, and that many
	// bytes come back with the status.  A block that produced more (the reference bounds it by the caller's Buffer
	// only) fetches the rest afterwards.
	const uint32_t spec = expect_out && expect_out <= (256u << 10) ? (expect_out < dst_cap ? expect_out : dst_cap) : 0u;
	if (spec && hash_content) {
		xxh32_stream_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<XxhState *>(s->d_meta + META_XXH), s->d_win + s->cursor, 0,
							       nullptr, d_st);
		ctx->launches++;
		CK(cudaGetLastError());
	}
	CK(cudaMemcpyAsync(s->h_meta + META_STATUS, s->d_meta + META_STATUS, sizeof(lz4b200_blk_status),
			   cudaMemcpyDeviceToHost, ctx->stream));
	if (spec) CK(cudaMemcpyAsync(host_dst, s->d_win + s->cursor, spec, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	*status = *reinterpret_cast<lz4b200_blk_status *>(s->h_meta + META_STATUS);
	if (status->code != LZ4B200_ST_OK) return LZ4B200_OK;
	const uint32_t n = status->out_len;
	if (n > spec) {
		if (hash_content && !spec) {
			xxh32_stream_kernel<<<1, 32, 0, ctx->stream>>>(
				reinterpret_cast<XxhState *>(s->d_meta + META_XXH), s->d_win + s->cursor, n, nullptr);
			ctx->launches++;
			CK(cudaGetLastError());
		}
		CK(cudaMemcpyAsync(host_dst + spec, s->d_win + s->cursor + spec, n - spec, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	s->cursor += n;
	s->frame_pos += n;
	return LZ4B200_OK;
}

int lz4b200_stream_block(lz4b200_stream *s, const uint8_t *host_src, uint32_t src_len, uint32_t flags,
			 int hash_content, uint8_t *host_dst, uint32_t dst_cap,
			 lz4b200_blk_status *status)
{
	return lz4b200_stream_block2(s, host_src, src_len, flags, hash_content, host_dst, dst_cap, 0, status);
}

int lz4b200_stream_adopt(lz4b200_stream *s, const uint8_t *dev_bytes, uint32_t n, int hash_content)
{
	if (!s || (!dev_bytes && n)) return LZ4B200_ERR_ARG;
	lz4b200_ctx *ctx = s->ctx;
	if (n > s->max_block) return LZ4B200_ERR_ARG;
	if (n == 0) return LZ4B200_OK;
	// keep the last 64 KiB in front of the cursor; compact when the window is exhausted
	if (s->cursor + n > s->win_size) {
		CK(cudaMemcpyAsync(s->d_win, s->d_win + s->cursor - HISTORY, HISTORY, cudaMemcpyDeviceToDevice,
				   ctx->stream));
		s->cursor = HISTORY;
	}
	CK(cudaMemcpyAsync(s->d_win + s->cursor, dev_bytes, n, cudaMemcpyDeviceToDevice, ctx->stream));
	if (hash_content) {
		xxh32_stream_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<XxhState *>(s->d_meta + META_XXH),
							       s->d_win + s->cursor, n, nullptr);
		ctx->launches++;
		CK(cudaGetLastError());
	}
	s->cursor += n;
	s->frame_pos += n;
	return LZ4B200_OK;
}

int lz4b200_stream_adopt_list(lz4b200_stream *s, const uint8_t *dev_base, uint32_t n_pieces, const uint32_t *offsets,
			      const uint32_t *lengths, int hash_content)
{
	if (!s || n_pieces > 255 || (n_pieces && (!dev_base || !offsets || !lengths))) return LZ4B200_ERR_ARG;
	if (n_pieces == 0) return LZ4B200_OK;
	lz4b200_ctx *ctx = s->ctx;
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_pieces; i++) {
		if (lengths[i] > s->max_block) return LZ4B200_ERR_ARG;
		total += lengths[i];
	}
	if (total == 0) return LZ4B200_OK;
	if (hash_content) {
		PieceList pl;
		pl.n = n_pieces;
		memcpy(pl.off, offsets, sizeof(uint32_t) * n_pieces);
		memcpy(pl.len, lengths, sizeof(uint32_t) * n_pieces);
		xxh32_stream_list_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<XxhState *>(s->d_meta + META_XXH), dev_base, pl);
		ctx->launches++;
		CK(cudaGetLastError());
	}
	if (total >= HISTORY) {
		// only the last 64 KiB can ever be referenced: they become the window
		uint64_t need = HISTORY;
		for (uint32_t i = n_pieces; i-- > 0 && need;) {
			const uint64_t take = lengths[i] < need ? lengths[i] : need;
			if (take) CK(cudaMemcpyAsync(s->d_win + need - take, dev_base + offsets[i] + lengths[i] - take, take, cudaMemcpyDeviceToDevice, ctx->stream));
			need -= take;
		}
		s->cursor = HISTORY;
	} else {
		for (uint32_t i = 0; i < n_pieces; i++) {
			const uint32_t n = lengths[i];
			if (!n) continue;
			if (s->cursor + n > s->win_size) {
				CK(cudaMemcpyAsync(s->d_win, s->d_win + s->cursor - HISTORY, HISTORY, cudaMemcpyDeviceToDevice, ctx->stream));
				s->cursor = HISTORY;
			}
			CK(cudaMemcpyAsync(s->d_win + s->cursor, dev_base + offsets[i], n, cudaMemcpyDeviceToDevice, ctx->stream));
			s->cursor += n;
		}
	}
	s->frame_pos += total;
	return LZ4B200_OK;
}

int lz4b200_stream_digest(lz4b200_stream *s, uint32_t *xxh32)
{
	if (!s || !xxh32) return LZ4B200_ERR_ARG;
	lz4b200_ctx *ctx = s->ctx;
	xxh32_stream_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<XxhState *>(s->d_meta + META_XXH),
						       s->d_win, 0, reinterpret_cast<uint32_t *>(s->d_meta + META_DIGEST));
	ctx->launches++;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(s->h_meta + META_DIGEST, s->d_meta + META_DIGEST, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	*xxh32 = *reinterpret_cast<uint32_t *>(s->h_meta + META_DIGEST);
	return LZ4B200_OK;
}

}  // extern "C"
