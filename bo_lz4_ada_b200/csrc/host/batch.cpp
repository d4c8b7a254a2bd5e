// batch.cpp -- the batched device entry point: block-table builder + batch scheduler.
//
// Host stage (lz4ada_batch_plan): every stream is walked by the same state machine that serves
// Update (walker.cpp), with a BlockEngine that records blocks instead of decoding them.  That
// gives header / size-word / Single_Frame errors exactly where the reference raises them
// (lib/lz4ada.adb:155-361, 525-585) and a table {compressed offset, size, stored flag, checksum
// flag} per block.  Nothing is decoded on the host.
//
// Device stage (lz4ada_batch_run): K1 over all independent blocks (a lane or a warp per block,
// chosen from the shape of the batch), K4 over linked frames (a CTA per chain), K3 over the frames
// that carry a content checksum -- no host round trip in between -- then one D2H of the status /
// digest arrays, folded in stream order into the first exception the reference would have raised
// (SURVEY.md Appendix A "ordering").
//
// Output placement: the frame format has no per-block decompressed size.  Blocks are placed
// optimistically at frame_base + i * block_max (true for every encoder that fills its blocks);
// the last block of a frame that is followed by another frame of the same stream is sized first
// by K5.  A stream that violates the assumption (short interior block, match into the previous
// block of an "independent" frame) is decoded again as one chain with exact running placement.
#include <thread>
#include <cstdlib>
#include <algorithm>
#include <memory>
#include <mutex>

#include "common.hpp"

namespace lz4ada {

struct FramePlan {
	Format format = Format::TBD;
	bool independent = true;
	bool has_cchk = false;     // FLG bit 2
	bool cchk_seen = false;    // the 4 checksum bytes were present in the stream
	uint32_t cchk_declared = 0;
	bool has_csize = false;
	uint64_t csize = 0;
	bool ended = false;        // end mark processed
	uint32_t first_block = 0, n_blocks = 0;
	uint32_t block_max = 0;
	uint64_t dst_off = 0;      // frame base in the batch output
	bool chained = false;      // linked frame: decoded by K4 as one chain
	bool solo = false;         // independent frame with big blocks: every block is its own K4 chain (a CTA each)
	uint32_t hash_slot = 0xffffffffu;
};

struct ItemPlan {
	uint64_t src_off = 0, src_len = 0, dst_off = 0, dst_cap = 0;
	bool user_placed = false;
	uint32_t first_frame = 0, n_frames = 0;   // data frames (legacy / modern)
	uint32_t frames_seen = 0;                 // including skippable
	uint32_t first_block = 0, n_blocks = 0;
	Raised host_error;
	int eof = LZ4ADA_EOF_NO;
	bool slow = false;
	bool overflowed = false;   // a block outgrew the frame's block maximum (the reference only bounds it by its Buffer)
	int min_buffer = 0;        // Min_Buffer_Size of the Init call the stream is decoded under (lib/lz4ada.adb:54, :119)
	// outcome of the last run
	Raised error;
	uint64_t out_len = 0;
};

class PlanEngine : public BlockEngine {
public:
	std::vector<FramePlan> *frames = nullptr;
	std::vector<lz4b200_blk_desc> *descs = nullptr;
	ItemPlan *item = nullptr;
	uint64_t call_pos = 0;   // stream offset of the input handed to the current update call
	int cur = -1;

	Raised new_frame(Walker &) override
	{
		cur = -1;
		return ok();
	}
	void frame_started(Walker &w) override
	{
		item->frames_seen++;
		if (w.m.format != Format::Legacy && w.m.format != Format::Modern) return;
		FramePlan f;
		f.format = w.m.format;
		f.independent = w.m.block_independent;
		f.has_cchk = w.m.content_checksum_length != 0;
		f.has_csize = w.m.has_content_size;
		f.csize = w.m.has_content_size ? w.m.size_remaining : 0;
		f.block_max = uint32_t(w.m.frame_block_max);
		f.first_block = uint32_t(descs->size());
		cur = int(frames->size());
		frames->push_back(f);
		item->n_frames++;
	}
	Raised block(Walker &w, const uint8_t *, int blk_len, uint8_t *, int, int &of, int &ol) override
	{
		of = 1;
		ol = 0;
		if (cur < 0) return err_assertion("block outside of a frame");
		const int raw_len = blk_len - w.m.block_checksum_length;
		lz4b200_blk_desc d;
		memset(&d, 0, sizeof d);
		d.src_off = item->src_off + call_pos + uint64_t(w.block_end_consumed) - uint64_t(blk_len);
		d.src_len = uint32_t(raw_len);
		d.flags = (w.m.is_compressed ? 0u : LZ4B200_BLK_STORED) |
			  (w.m.block_checksum_length ? LZ4B200_BLK_HAS_CHECKSUM : 0u);
		descs->push_back(d);
		(*frames)[size_t(cur)].n_blocks++;
		item->n_blocks++;
		return ok();
	}
	Raised content_checksum(Walker &, uint32_t declared) override
	{
		if (cur >= 0) {
			(*frames)[size_t(cur)].cchk_seen = true;
			(*frames)[size_t(cur)].cchk_declared = declared;
		}
		return ok();   // compared after the device run
	}
	Raised frame_ended(Walker &) override
	{
		if (cur >= 0) (*frames)[size_t(cur)].ended = true;
		return ok();   // content-size-left is checked after the device run
	}
};

static uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// Host threads one plan may use for its walk.  Eight is plenty for one process; processes (or the threads of the
// multi-GPU call) that share a box share its cores: LZ4ADA_PLAN_THREADS overrides, LOCAL_WORLD_SIZE (torchrun) divides.
static thread_local uint32_t g_plan_share = 1;   // set by lz4ada_batch_decompress_multi for its worker threads
static uint32_t plan_threads()
{
	const uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
	if (const char *e = getenv("LZ4ADA_PLAN_THREADS")) {
		const int v = atoi(e);
		if (v >= 1) return std::min<uint32_t>(uint32_t(v), 64u);
	}
	uint32_t share = g_plan_share;
	if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
		const int v = atoi(e);
		if (v > 1) share = std::max<uint32_t>(share, uint32_t(v));
	}
	return std::max(1u, std::min<uint32_t>(8u, hw / share));
}

}  // namespace lz4ada

using namespace lz4ada;

struct lz4ada_batch {
	lz4b200_ctx *ctx = nullptr;
	int reservation = LZ4ADA_FOR_ALL;
	uint64_t src_bytes = 0;
	uint64_t span_lo = 0, span_hi = 0;   // the part of the source buffer the streams of this batch occupy
	std::vector<ItemPlan> items;
	std::vector<FramePlan> frames;
	std::vector<lz4b200_blk_desc> descs;
	std::vector<lz4b200_chain> chains;
	std::vector<lz4b200_frame_blocks> hash_frames;   // frames whose content checksum K3 computes
	std::vector<uint32_t> k2_idx;                    // stored blocks, copied by K2 (lz4b200_copy_stored) instead of K1
	std::vector<lz4b200_hash_span> k2_spans;         // ... their payloads, for the block checksums
	bool k2_checksums = false;
	uint32_t *d_k2_idx = nullptr, *d_k2_scratch = nullptr;
	lz4b200_hash_span *d_k2_spans = nullptr;
	size_t cap_k2 = 0;
	bool exact_sizing = false;                       // lz4ada_batch_exact_sizing: K5 sizes every block first
	int64_t heavy_blocks = -1;                       // blocks that take long to decode (K1 kernel choice), -1 = not counted yet
	const char *k1_name = "";                        // the K1 kernel the last lz4ada_batch_run launched
	std::vector<uint32_t> presize;                   // blocks K5 sizes before placement
	std::vector<uint32_t> presize_max;               // block maximum of the frame each of them belongs to
	std::vector<uint32_t> sized;                     // their sizes once K5 has run
	bool have_sized = false;
	uint64_t out_bytes = 0;
	uint64_t dst_capacity = 0;   // what the caller's output buffer really holds (lz4ada_batch_set_output_capacity), 0 = out_bytes
	uint64_t spill_cursor = 0;   // streams that outgrow their region are decoded again behind out_bytes
	bool placed = false;
	bool tables_uploaded = false;
	// device tables
	lz4b200_blk_desc *d_desc = nullptr;
	lz4b200_blk_status *d_status = nullptr;
	lz4b200_chain *d_chains = nullptr;         // the plan's chains (linked frames), uploaded once
	lz4b200_chain *d_retry_chains = nullptr;   // chains of the streams a run decodes again (run_slow_items)
	size_t cap_retry_chains = 0;
	lz4b200_frame_blocks *d_hash_frames = nullptr;
	uint32_t *d_digest = nullptr;   // [n_hash] digests then [n_hash] valid flags
	size_t cap_chains = 0;
	// pinned host mirrors
	lz4b200_blk_status *h_status = nullptr;
	uint32_t *h_digest = nullptr;
	// traffic and kernel times of the last run
	uint64_t t_comp = 0, t_out = 0, t_reread = 0;
	void *ev[4] = {nullptr, nullptr, nullptr, nullptr};
	float kernel_ms[3] = {0, 0, 0};
	std::string device_error;

	~lz4ada_batch()
	{
		if (!ctx) return;   // never reached the device
		lz4b200_sync(ctx);
		if (d_desc) lz4b200_free(ctx, d_desc);
		if (d_status) lz4b200_free(ctx, d_status);
		if (d_chains) lz4b200_free(ctx, d_chains);
		if (d_retry_chains) lz4b200_free(ctx, d_retry_chains);
		if (d_k2_idx) lz4b200_free(ctx, d_k2_idx);
		if (d_k2_spans) lz4b200_free(ctx, d_k2_spans);
		if (d_k2_scratch) lz4b200_free(ctx, d_k2_scratch);
		if (d_hash_frames) lz4b200_free(ctx, d_hash_frames);
		if (d_digest) lz4b200_free(ctx, d_digest);
		if (h_status) lz4b200_free_host(ctx, h_status);
		if (h_digest) lz4b200_free_host(ctx, h_digest);
		for (void *e : ev)
			if (e) lz4b200_event_destroy(ctx, e);
		lz4b200_destroy(ctx);   // the batch's reference
	}
};

namespace {

// Optimistic placement.  `sized` = out_len of pre-sized blocks (indexed like b->presize) or null.
void place(lz4ada_batch *b, const std::vector<uint32_t> *sized_len)
{
	std::vector<int64_t> known(b->descs.size(), -1);
	if (sized_len)
		for (size_t i = 0; i < b->presize.size(); i++)
			if ((*sized_len)[i] != 0xffffffffu) known[b->presize[i]] = (*sized_len)[i];   // 0xffffffff: K5 gave up on it
	b->chains.clear();
	b->hash_frames.clear();
	b->k2_idx.clear();
	b->k2_spans.clear();
	b->k2_checksums = false;
	// Big independent blocks (>= 64 KiB of compressed bytes: >= 10^4 sequences in series) are placed as chains of one
	// block while the batch holds few of them: K1 gives such a block one warp, which walks it at ~100 MB/s whatever
	// the kernel generation; the chain kernel K7 gives it a CTA that parses and copies it without a serial walk
	// (kernels_k7.cuh: 250 - 400 MB/s per stream) but spends 2.3x the instructions, so its whole-chip rate stops at
	// 40 - 60 GB/s where K1 v4 with thousands of streams reaches 200 - 300.  Measured: 341 text blocks of 4 MiB 24.5 ms
	// as chains against 41.8 ms in K1; ~550 blocks of 250 - 450 KiB 5.8 ms against ~4 ms.  Chains up to 400 big blocks
	// per batch, K1 beyond.  Highly compressible big blocks (zero pages, RLE: a few KiB of
	// compressed bytes, some giant matches) stay in K1 either way, whose warp-wide copies are what they need.
	// LZ4B200_SOLO=0 turns the placement off, LZ4B200_SOLO_MAX=n moves the limit (A/B switches).
	static const int solo_max = [] {
		const char *e = getenv("LZ4B200_SOLO");
		if (e && e[0] == '0') return 0;
		const char *m = getenv("LZ4B200_SOLO_MAX");
		return m ? atoi(m) : 400;
	}();
	bool solo_on = false;
	if (solo_max > 0) {
		size_t big = 0;
		for (const FramePlan &fp : b->frames) {
			if (!fp.independent && fp.n_blocks > 1) continue;
			for (uint32_t i = 0; i < fp.n_blocks; i++) {
				const lz4b200_blk_desc &d = b->descs[fp.first_block + i];
				if (!(d.flags & LZ4B200_BLK_STORED) && d.src_len >= 65536) big++;
			}
		}
		solo_on = big > 0 && big <= size_t(solo_max);
	}
	uint64_t cursor = 0;
	for (ItemPlan &it : b->items) {
		it.slow = false;
		uint64_t upper = 0;
		for (uint32_t f = 0; f < it.n_frames; f++) {
			const FramePlan &fp = b->frames[it.first_frame + f];
			upper += uint64_t(fp.n_blocks) * fp.block_max;
		}
		if (!it.user_placed) {
			it.dst_off = align_up(cursor, 256);
			it.dst_cap = upper;
		}
		const uint64_t item_end = it.dst_off + it.dst_cap;
		uint64_t pos = it.dst_off;
		for (uint32_t f = 0; f < it.n_frames; f++) {
			FramePlan &fp = b->frames[it.first_frame + f];
			fp.dst_off = pos;
			fp.chained = !fp.independent && fp.n_blocks > 1;
			// a big block is ~10^5 sequences in series: give it the chain kernel (one CTA)
			fp.solo = solo_on && !fp.chained;
			fp.hash_slot = 0xffffffffu;
			// blocks sit one block-maximum apart (the frame format has no per-block decompressed size) -- unless K5
			// has sized every block in front of this one (exact-sizing mode): then they sit back to back
			uint64_t run = 0;
			bool run_known = true;
			for (uint32_t i = 0; i < fp.n_blocks; i++) {
				lz4b200_blk_desc &d = b->descs[fp.first_block + i];
				const uint64_t off_i = run_known && !fp.chained ? run : uint64_t(i) * fp.block_max;
				d.dst_off = fp.dst_off + off_i;
				const uint64_t room = d.dst_off < item_end ? item_end - d.dst_off : 0;
				const int64_t ki = known[fp.first_block + i];
				const bool exact_i = run_known && !fp.chained && ki >= 0;
				d.dst_cap = uint32_t(std::min<uint64_t>(exact_i ? uint64_t(ki) : fp.block_max, room));
				if (!exact_i && room < fp.block_max && i + 1 < fp.n_blocks) it.slow = true;   // tight user buffer
				if (exact_i && room < uint64_t(ki)) it.slow = true;
				if (ki >= 0) run += uint64_t(ki);
				else run_known = false;
				const uint64_t hist = off_i;
				d.hist_avail = hist > 0xfffffffeull ? 0xffffffffu : uint32_t(hist);
				d.flags &= ~(LZ4B200_BLK_CHAINED | LZ4B200_BLK_FIRST_OF_FRAME | LZ4B200_BLK_SOLO | LZ4B200_BLK_RING_CAP | LZ4B200_BLK_K2);
				if (fp.chained) d.flags |= LZ4B200_BLK_CHAINED;
				if (i == 0) d.flags |= LZ4B200_BLK_FIRST_OF_FRAME;
				if (!fp.chained && (d.flags & LZ4B200_BLK_STORED) && !(d.flags & LZ4B200_BLK_HASH_ONLY)) {
					// stored blocks of independent frames: the wide copy K2, not a warp of K1
					d.flags |= LZ4B200_BLK_K2;
					b->k2_idx.push_back(fp.first_block + i);
					lz4b200_hash_span sp;
					sp.off = d.src_off;
					sp.len = d.src_len;
					b->k2_spans.push_back(sp);
					if (d.flags & LZ4B200_BLK_HAS_CHECKSUM) b->k2_checksums = true;
				}
				if (fp.solo && !(d.flags & LZ4B200_BLK_STORED) && d.src_len >= 65536) {
					d.flags |= LZ4B200_BLK_CHAINED | LZ4B200_BLK_FIRST_OF_FRAME | LZ4B200_BLK_SOLO;
					lz4b200_chain c;
					c.first_block = fp.first_block + i;
					c.n_blocks = 1;
					c.dst_off = d.dst_off;
					c.dst_cap = d.dst_cap;
					b->chains.push_back(c);
				}
			}
			if (fp.chained) {
				lz4b200_chain c;
				c.first_block = fp.first_block;
				c.n_blocks = fp.n_blocks;
				c.dst_off = fp.dst_off;
				c.dst_cap = fp.dst_off < item_end ? std::min<uint64_t>(uint64_t(fp.n_blocks) * fp.block_max,
										      item_end - fp.dst_off)
								  : 0;
				b->chains.push_back(c);
			}
			if (fp.has_cchk && fp.cchk_seen && fp.n_blocks) {   // (a frame without blocks has the checksum of no bytes: host)
				fp.hash_slot = uint32_t(b->hash_frames.size());
				lz4b200_frame_blocks hb;
				hb.first_block = fp.first_block;
				hb.n_blocks = fp.n_blocks;
				b->hash_frames.push_back(hb);
			}
			// where the next frame of this stream starts: interior blocks full, last block sized by K5
			uint64_t fsize = 0;
			if (fp.n_blocks) {
				const uint32_t last = fp.first_block + fp.n_blocks - 1;
				const int64_t k = known[last];
				if (run_known && !fp.chained) fsize = run;   // every block sized: the frame is exactly this long
				else fsize = uint64_t(fp.n_blocks - 1) * fp.block_max + (k >= 0 ? uint64_t(k) : fp.block_max);
			}
			pos += fsize;
		}
		if (!it.user_placed) cursor = it.dst_off + it.dst_cap;
		else cursor = std::max(cursor, item_end);
	}
	b->out_bytes = std::max(b->out_bytes, cursor);   // never shrinks: callers allocate from the first answer
	b->spill_cursor = b->out_bytes;
	b->placed = true;
}

// Blocks of [b0, b1) that suit the lane-per-block K1: compressed, of some size, up to 256 KiB of output, not RLE-like.
uint64_t heavy_blocks(const lz4ada_batch *b, size_t b0, size_t b1)
{
	uint64_t n = 0;
	for (size_t i = b0; i < b1; i++) {
		const lz4b200_blk_desc &d = b->descs[i];
		if (!(d.flags & (LZ4B200_BLK_STORED | LZ4B200_BLK_CHAINED | LZ4B200_BLK_HASH_ONLY)) && d.src_len >= 4096 && d.dst_cap <= 262144 &&
		    uint64_t(d.src_len) * 16 >= d.dst_cap)
			n++;
	}
	return n;
}

// K1 over blocks [b0, b1): the block-count rule of lz4b200_decode_blocks, overridden towards the warp-per-block
// kernel when too few of the blocks suit the lane-per-block one.
int launch_k1(lz4ada_batch *b, const uint8_t *src_dev, uint8_t *dst_dev, size_t b0, size_t b1, uint64_t heavy)
{
	lz4b200_ctx *ctx = b->ctx;
	const int saved_tuning = lz4b200_get_tuning(ctx);
	const bool force_v4 = saved_tuning == 0 && heavy < 20000;
	if (force_v4) lz4b200_set_tuning(ctx, 40);
	b->k1_name = lz4b200_k1_kernel_name(ctx, uint32_t(b1 - b0));
	const int rc = lz4b200_decode_blocks(ctx, src_dev, dst_dev, uint32_t(b1 - b0), b->d_desc + b0, b->d_status + b0);
	if (force_v4) lz4b200_set_tuning(ctx, saved_tuning);
	return rc;
}

// K2 over the stored blocks among [b0, b1) (the index list is in block order).
int launch_k2(lz4ada_batch *b, const uint8_t *src_dev, uint8_t *dst_dev, size_t b0, size_t b1)
{
	const auto lo = std::lower_bound(b->k2_idx.begin(), b->k2_idx.end(), uint32_t(b0));
	const auto hi = std::lower_bound(b->k2_idx.begin(), b->k2_idx.end(), uint32_t(b1));
	if (lo == hi) return LZ4B200_OK;
	const size_t k0 = size_t(lo - b->k2_idx.begin()), n = size_t(hi - lo);
	uint32_t max_len = 0;
	for (size_t k = k0; k < k0 + n; k++) max_len = std::max(max_len, uint32_t(b->k2_spans[k].len));
	return lz4b200_copy_stored(b->ctx, src_dev, dst_dev, uint32_t(n), b->d_k2_idx + k0, max_len, b->d_desc, b->d_status,
				   b->k2_checksums ? b->d_k2_spans + k0 : nullptr, b->k2_checksums ? b->d_k2_scratch + k0 : nullptr);
}

Raised device_fail(lz4ada_batch *b)
{
	b->device_error = lz4b200_last_error(b->ctx);
	return err_device(b->device_error.c_str());
}

// Fold the statuses of one stream in stream order.  exact = the stream was decoded as one chain
// (placement is exact by construction; frame bases are recomputed here).
// digest_of(frame) must deliver (valid, value).
// Returns true when the stream has to be decoded again as one chain (placement assumption broken).
template <class DigestFn>
bool fold_item(lz4ada_batch *b, ItemPlan &it, bool exact, DigestFn digest_of)
{
	if (!exact && it.slow) return true;   // placement already known to be unusable (tight caller buffer)
	if (!exact) it.overflowed = false;
	it.error = Raised();
	it.out_len = 0;
	bool slow = false;
	uint64_t pos = it.dst_off;
	for (uint32_t f = 0; f < it.n_frames && !it.error && !slow; f++) {
		FramePlan &fp = b->frames[it.first_frame + f];
		if (exact) fp.dst_off = pos;
		else if (fp.dst_off != pos) { slow = true; break; }
		uint64_t fpos = 0;
		uint64_t remaining = fp.csize;
		uint32_t ring = 0;   // the reference's Output_Pos (exact runs bound a block by the caller's Buffer, RING_CAP)
		for (uint32_t i = 0; i < fp.n_blocks; i++) {
			const uint32_t bi = fp.first_block + i;
			const lz4b200_blk_status &st = b->h_status[bi];
			const lz4b200_blk_desc &d = b->descs[bi];
			if (ring >= uint32_t(kHistorySize)) ring = 0;
			if (st.code == LZ4B200_ST_NEEDS_HISTORY || st.code == LZ4B200_ST_NOT_RUN || st.code == 0xffffffffu) { slow = true; break; }
			if (!exact && !fp.chained && d.dst_off != fp.dst_off + fpos) { slow = true; break; }
			if (fp.has_csize && st.code != LZ4B200_ST_BLOCK_CHECKSUM) {
				const uint64_t produced = st.code == LZ4B200_ST_OK ? st.out_len : st.err_pos;
				if (remaining < produced) { it.error = err_content_size_exceeded(); break; }
			}
			if (st.code != LZ4B200_ST_OK) {
				if (st.code == LZ4B200_ST_OUTPUT_OVERFLOW && !exact) {
					// The slot (one block maximum, or less in a tight region) was a placement assumption: the reference
					// bounds a block's output only by the caller's Buffer (lib/lz4ada.adb:54, 813-820).  Decode the
					// stream again as a chain under that bound.
					it.overflowed = true;
					slow = true;
					break;
				}
				int shown = int(std::min<uint64_t>(d.dst_cap, 0x7fffffff));
				if (st.code == LZ4B200_ST_OUTPUT_OVERFLOW && (d.flags & LZ4B200_BLK_RING_CAP)) {
					// which bound was hit: the Buffer of the reference's caller, or the room this stream has in the batch output
					const uint64_t room = it.dst_off + it.dst_cap - (pos + fpos);
					const uint64_t by_ring = uint64_t(it.min_buffer) > ring ? uint64_t(it.min_buffer) - ring : 0;
					shown = by_ring <= room ? it.min_buffer : int(std::min<uint64_t>(it.dst_cap, 0x7fffffff));
				}
				it.error = status_to_raised(st, shown);
				break;
			}
			remaining -= st.out_len;
			fpos += st.out_len;
			ring += st.out_len;
		}
		// the reference hands out every block before the failing one (one block per Update)
		pos += fpos;
		it.out_len += fpos;
		if (it.error || slow) break;
		if (fp.has_cchk && fp.cchk_seen) {
			uint32_t value = 0;
			const uint8_t none = 0;
			if (fp.n_blocks == 0) value = Xxh32Host::hash(&none, 0);   // XXHash32.Final of nothing, lib/lz4ada.adb:993-1017
			else if (!digest_of(fp, value)) { slow = true; break; }
			if (value != fp.cchk_declared) { it.error = err_content_checksum(value, fp.cchk_declared); break; }
		}
		if (fp.ended && fp.has_csize && remaining != 0) { it.error = err_content_size_left(remaining); break; }
	}
	if (slow) return true;
	if (!it.error) it.error = it.host_error;
	return false;
}

// Streams whose blocks outgrew their slots: size every block with K5 and, where the stream no longer fits the
// region the plan gave it, move it behind the planned output (as far as the caller's buffer reaches).
Raised make_room(lz4ada_batch *b, const uint8_t *src_dev)
{
	uint32_t lo = 0xffffffffu, hi = 0;
	for (const ItemPlan &it : b->items)
		if (it.slow && it.overflowed && it.n_blocks && !it.user_placed) {
			lo = std::min(lo, it.first_block);
			hi = std::max(hi, it.first_block + it.n_blocks);
		}
	if (lo >= hi) return ok();
	// (one launch over the block range that spans them: K5 reads the compressed bytes only, a thread per block)
	const uint32_t n = hi - lo;
	lz4b200_blk_status *d_ps = nullptr;
	std::vector<lz4b200_blk_status> ps(n);
	if (lz4b200_alloc(b->ctx, sizeof(lz4b200_blk_status) * n, reinterpret_cast<void **>(&d_ps)) != LZ4B200_OK) return device_fail(b);
	const bool bad = lz4b200_size_blocks(b->ctx, src_dev, n, b->d_desc + lo, d_ps) != LZ4B200_OK ||
			 lz4b200_d2h(b->ctx, ps.data(), d_ps, sizeof(lz4b200_blk_status) * n) != LZ4B200_OK ||
			 lz4b200_sync(b->ctx) != LZ4B200_OK;
	lz4b200_free(b->ctx, d_ps);
	if (bad) return device_fail(b);
	const uint64_t capacity = std::max(b->dst_capacity, b->out_bytes);
	for (ItemPlan &it : b->items) {
		if (!(it.slow && it.overflowed && it.n_blocks && !it.user_placed)) continue;
		uint64_t total = 0;
		for (uint32_t i = 0; i < it.n_blocks; i++) total += ps[it.first_block + i - lo].out_len;   // up to the first error, if any
		if (total <= it.dst_cap) continue;
		const uint64_t at = align_up(b->spill_cursor, 256);
		if (at + total > capacity) continue;   // no room: the chain reports the overflow against the region it has
		it.dst_off = at;
		it.dst_cap = total;
		b->spill_cursor = at + total;
	}
	return ok();
}

// Decode the flagged streams again, each as one chain with exact running placement.  In a chain a block is
// bounded the way the reference bounds it -- by what is left of the caller's Buffer behind the ring cursor
// (LZ4B200_BLK_RING_CAP), not by the frame's block maximum.
Raised run_slow_items(lz4ada_batch *b, const uint8_t *src_dev, uint8_t *dst_dev)
{
	if (Raised r = make_room(b, src_dev)) return r;
	std::vector<lz4b200_chain> chains;
	for (ItemPlan &it : b->items) {
		if (!it.slow || it.n_blocks == 0) continue;
		for (uint32_t f = 0; f < it.n_frames; f++) {
			FramePlan &fp = b->frames[it.first_frame + f];
			for (uint32_t i = 0; i < fp.n_blocks; i++) {
				lz4b200_blk_desc &d = b->descs[fp.first_block + i];
				d.flags &= ~(LZ4B200_BLK_FIRST_OF_FRAME | LZ4B200_BLK_SOLO | LZ4B200_BLK_K2);
				d.flags |= LZ4B200_BLK_CHAINED | LZ4B200_BLK_RING_CAP;
				if (i == 0) d.flags |= LZ4B200_BLK_FIRST_OF_FRAME;
				d.dst_cap = uint32_t(it.min_buffer);
			}
		}
		lz4b200_chain c;
		c.first_block = it.first_block;
		c.n_blocks = it.n_blocks;
		c.dst_off = it.dst_off;
		c.dst_cap = it.dst_cap;
		chains.push_back(c);
		if (lz4b200_h2d(b->ctx, b->d_desc + it.first_block, b->descs.data() + it.first_block,
				sizeof(lz4b200_blk_desc) * it.n_blocks) != LZ4B200_OK)
			return device_fail(b);
	}
	if (chains.empty()) return ok();
	// (a buffer of their own: b->d_chains keeps the plan's chains for the next run)
	if (chains.size() > b->cap_retry_chains) {
		if (b->d_retry_chains) lz4b200_free(b->ctx, b->d_retry_chains);
		b->d_retry_chains = nullptr;
		b->cap_retry_chains = 0;
		if (lz4b200_alloc(b->ctx, sizeof(lz4b200_chain) * chains.size(), reinterpret_cast<void **>(&b->d_retry_chains)) != LZ4B200_OK)
			return device_fail(b);
		b->cap_retry_chains = chains.size();
	}
	if (lz4b200_h2d(b->ctx, b->d_retry_chains, chains.data(), sizeof(lz4b200_chain) * chains.size()) != LZ4B200_OK ||
	    lz4b200_decode_linked(b->ctx, src_dev, dst_dev, uint32_t(chains.size()), b->d_retry_chains, b->d_desc, b->d_status) != LZ4B200_OK ||
	    lz4b200_d2h(b->ctx, b->h_status, b->d_status, sizeof(lz4b200_blk_status) * b->descs.size()) != LZ4B200_OK ||
	    lz4b200_sync(b->ctx) != LZ4B200_OK)
		return device_fail(b);
	// content checksums of the re-placed frames: spans are known only now
	std::vector<lz4b200_hash_span> spans;
	std::vector<FramePlan *> span_frames;
	for (ItemPlan &it : b->items) {
		if (!it.slow) continue;
		uint64_t pos = it.dst_off;
		for (uint32_t f = 0; f < it.n_frames; f++) {
			FramePlan &fp = b->frames[it.first_frame + f];
			uint64_t len = 0;
			bool okay = true;
			for (uint32_t i = 0; i < fp.n_blocks; i++) {
				const lz4b200_blk_status &st = b->h_status[fp.first_block + i];
				if (st.code != LZ4B200_ST_OK) { okay = false; break; }
				len += st.out_len;
			}
			if (!okay) break;
			if (fp.has_cchk && fp.cchk_seen) {
				lz4b200_hash_span s;
				s.off = pos;
				s.len = len;
				spans.push_back(s);
				span_frames.push_back(&fp);
			}
			pos += len;
		}
	}
	std::vector<uint32_t> values(spans.size());
	if (!spans.empty()) {
		lz4b200_hash_span *d_spans = nullptr;
		uint32_t *d_out = nullptr;
		if (lz4b200_alloc(b->ctx, sizeof(lz4b200_hash_span) * spans.size(), reinterpret_cast<void **>(&d_spans)) != LZ4B200_OK ||
		    lz4b200_alloc(b->ctx, 4 * spans.size(), reinterpret_cast<void **>(&d_out)) != LZ4B200_OK)
			return device_fail(b);
		bool bad = lz4b200_h2d(b->ctx, d_spans, spans.data(), sizeof(lz4b200_hash_span) * spans.size()) != LZ4B200_OK ||
			   lz4b200_xxh32_spans(b->ctx, dst_dev, uint32_t(spans.size()), d_spans, d_out) != LZ4B200_OK ||
			   lz4b200_d2h(b->ctx, values.data(), d_out, 4 * spans.size()) != LZ4B200_OK ||
			   lz4b200_sync(b->ctx) != LZ4B200_OK;
		lz4b200_free(b->ctx, d_spans);
		lz4b200_free(b->ctx, d_out);
		if (bad) return device_fail(b);
	}
	for (ItemPlan &it : b->items) {
		if (!it.slow) continue;
		const bool again = fold_item(b, it, true, [&](const FramePlan &fp, uint32_t &value) {
			for (size_t k = 0; k < span_frames.size(); k++)
				if (span_frames[k] == &fp) { value = values[k]; return true; }
			return false;
		});
		if (again) {   // cannot happen: the chain path has no soft failures
			it.error = err_device("chain decode reported an unexpected soft status");
		}
	}
	return ok();
}

}  // namespace

extern "C" {

int lz4ada_batch_plan(lz4b200_ctx *ctx, const uint8_t *src_host, uint64_t src_bytes, uint32_t n_items,
		      const lz4ada_batch_item *items, int reservation, lz4ada_batch **out)
{
	if (!out || (!items && n_items) || reservation < LZ4ADA_SZ_64_KIB || reservation > LZ4ADA_SINGLE_FRAME)
		return LZ4ADA_ASSERTION_ERROR;
	*out = nullptr;
	// planning is pure host work; the device context is only needed from upload on
	std::unique_ptr<lz4ada_batch> b(new lz4ada_batch());
	if (ctx) lz4b200_retain(ctx);
	b->ctx = ctx;
	b->reservation = reservation;
	b->src_bytes = src_bytes;
	b->items.resize(n_items);
	b->span_lo = n_items ? ~0ull : 0;
	for (uint32_t k = 0; k < n_items; k++) {
		b->span_lo = std::min(b->span_lo, items[k].src_off);
		b->span_hi = std::max(b->span_hi, items[k].src_off + items[k].src_len);
	}
	if (b->span_hi < b->span_lo) b->span_lo = b->span_hi = 0;
	// A fixed reservation = Init (lib/lz4ada.adb:48-63); Use_First / Single_Frame = Init_With_Header on the whole
	// stream (:79-125), the call the reference's unlz4ada and its error-case test make
	const bool with_header = reservation > LZ4ADA_SZ_8_MIB;
	const int in_last = with_header ? 0 : block_size_of(reservation) + 4 + kBlockSizeBytes - 1;   // as Init, :60
	for (uint32_t k = 0; k < n_items; k++)
		if (items[k].src_off > src_bytes || items[k].src_len > src_bytes - items[k].src_off ||
		    items[k].dst_off + items[k].dst_cap < items[k].dst_off)   // a caller-placed region may not wrap
			return LZ4ADA_ASSERTION_ERROR;
	// Streams are independent: walk them on a few host threads (the walk touches one size word per block, spread
	// over the whole compressed buffer -- 10 ms for 4096 frames / 65536 blocks on one thread), each thread into its
	// own frame / block tables, which are then appended in stream order.
	struct Part {
		std::vector<FramePlan> frames;
		std::vector<lz4b200_blk_desc> descs;
	};
	const uint32_t n_parts = std::max(1u, std::min<uint32_t>(plan_threads(), n_items / 64u));
	std::vector<Part> parts(n_parts);
	auto walk_range = [&](uint32_t part) {
		Part &pt = parts[part];
		PlanEngine engine;
		engine.frames = &pt.frames;
		engine.descs = &pt.descs;
		const uint32_t k0 = uint32_t(uint64_t(n_items) * part / n_parts), k1 = uint32_t(uint64_t(n_items) * (part + 1) / n_parts);
		for (uint32_t k = k0; k < k1; k++) {
			ItemPlan &it = b->items[k];
			it.src_off = items[k].src_off;
			it.src_len = items[k].src_len;
			it.dst_off = items[k].dst_off;
			it.dst_cap = items[k].dst_cap;
			it.user_placed = items[k].dst_cap != 0;
			it.first_frame = uint32_t(pt.frames.size());   // part-relative until the tables are joined
			it.first_block = uint32_t(pt.descs.size());
			Meta m;
			m.reservation = reservation;
			const uint8_t *s = src_host + it.src_off;
			uint64_t pos = 0;
			int in_last_k = in_last;
			it.min_buffer = with_header ? 0 : block_size_of(reservation) + kHistorySize + 8;   // :54
			if (with_header) {
				if (it.src_len < 7) {   // Pre => Input'Length >= 7, lib/lz4ada.ads:243
					it.host_error = err_assertion("failed precondition from lz4ada.ads:243");
					continue;
				}
				int consumed = 0;
				Raised r = init_with_header_meta(s, int(std::min<uint64_t>(it.src_len, 0x40000000ull)), reservation, m, consumed,
								 in_last_k, it.min_buffer);
				if (r) {
					it.host_error = r;
					continue;
				}
				pos = uint64_t(consumed);
			}
			Walker w(m, in_last_k, &engine);
			engine.item = &it;
			engine.cur = -1;
			if (with_header) engine.frame_started(w);   // the first header was consumed by the Init call
			int idle = 0;
			while (pos < it.src_len) {
				const uint64_t left = it.src_len - pos;
				const int window = int(std::min<uint64_t>(left, 0x40000000ull));
				int consumed = 0, of = 1, ol = 0;
				engine.call_pos = pos;
				Raised r = w.update(s + pos, window, consumed, nullptr, 0, of, ol);
				if (r) {
					it.host_error = r;
					break;
				}
				pos += uint64_t(consumed);
				if (consumed == 0 && ++idle > 2) {
					it.host_error = err_assertion("No more data accepted but no exception signalled.");
					break;
				}
				if (consumed) idle = 0;
			}
			it.eof = w.is_end_of_frame();
		}
	};
	if (n_parts == 1) {
		walk_range(0);
	} else {
		std::vector<std::thread> workers;
		for (uint32_t p = 1; p < n_parts; p++) workers.emplace_back(walk_range, p);
		walk_range(0);
		for (std::thread &t : workers) t.join();
	}
	for (uint32_t part = 0; part < n_parts; part++) {
		Part &pt = parts[part];
		const uint32_t f0 = uint32_t(b->frames.size()), d0 = uint32_t(b->descs.size());
		const uint32_t k0 = uint32_t(uint64_t(n_items) * part / n_parts), k1 = uint32_t(uint64_t(n_items) * (part + 1) / n_parts);
		for (uint32_t k = k0; k < k1; k++) {
			b->items[k].first_frame += f0;
			b->items[k].first_block += d0;
		}
		for (FramePlan &fp : pt.frames) fp.first_block += d0;
		b->frames.insert(b->frames.end(), pt.frames.begin(), pt.frames.end());
		b->descs.insert(b->descs.end(), pt.descs.begin(), pt.descs.end());
	}
	// blocks whose size decides where the next frame of the same stream starts
	for (ItemPlan &it : b->items)
		for (uint32_t f = 0; f + 1 < it.n_frames; f++) {
			const FramePlan &fp = b->frames[it.first_frame + f];
			if (fp.n_blocks) {
				b->presize.push_back(fp.first_block + fp.n_blocks - 1);
				b->presize_max.push_back(fp.block_max);
			}
		}
	if (b->presize.empty()) place(b.get(), nullptr);
	else {
		// upper bound so that callers can allocate before the sizes are known
		place(b.get(), nullptr);
		b->placed = false;
	}
	*out = b.release();
	return LZ4ADA_OK;
}

int lz4ada_batch_block_desc(const lz4ada_batch *b, uint64_t index, lz4b200_blk_desc *out)
{
	if (!b || !out || index >= b->descs.size()) return LZ4ADA_ASSERTION_ERROR;
	*out = b->descs[index];
	return LZ4ADA_OK;
}

int lz4ada_batch_host_outcome(const lz4ada_batch *b, uint32_t item, int *exception, int *end_of_frame,
			      uint32_t *n_frames, uint32_t *n_blocks, char *message, size_t message_cap)
{
	if (!b || item >= b->items.size()) return LZ4ADA_ASSERTION_ERROR;
	const ItemPlan &it = b->items[item];
	if (exception) *exception = it.host_error.kind;
	if (end_of_frame) *end_of_frame = it.eof;
	if (n_frames) *n_frames = it.frames_seen;
	if (n_blocks) *n_blocks = it.n_blocks;
	if (message && message_cap) {
		const size_t n = std::min(it.host_error.text.size(), message_cap - 1);
		memcpy(message, it.host_error.text.data(), n);
		message[n] = 0;
	}
	return LZ4ADA_OK;
}

uint64_t lz4ada_batch_output_bytes(const lz4ada_batch *b) { return b ? b->out_bytes : 0; }
uint64_t lz4ada_batch_block_count(const lz4ada_batch *b) { return b ? b->descs.size() : 0; }

void lz4ada_batch_traffic(const lz4ada_batch *b, uint64_t *compressed_read, uint64_t *decompressed_written,
			  uint64_t *checksum_reread)
{
	if (compressed_read) *compressed_read = b ? b->t_comp : 0;
	if (decompressed_written) *decompressed_written = b ? b->t_out : 0;
	if (checksum_reread) *checksum_reread = b ? b->t_reread : 0;
}

int lz4ada_batch_set_output_capacity(lz4ada_batch *b, uint64_t bytes)
{
	if (!b) return LZ4ADA_ASSERTION_ERROR;
	b->dst_capacity = bytes;
	return LZ4ADA_OK;
}

int lz4ada_batch_exact_sizing(lz4ada_batch *b)
{
	if (!b || b->tables_uploaded) return LZ4ADA_ASSERTION_ERROR;
	b->exact_sizing = true;
	b->presize.clear();
	b->presize_max.clear();
	for (const FramePlan &fp : b->frames) {
		if (!(fp.independent || fp.n_blocks <= 1)) continue;   // linked frames are chains: placed exactly as they run
		for (uint32_t i = 0; i < fp.n_blocks; i++) {
			b->presize.push_back(fp.first_block + i);
			b->presize_max.push_back(fp.block_max);
		}
	}
	// the last block of a linked frame that is followed by another frame still decides that frame's base
	for (const ItemPlan &it : b->items)
		for (uint32_t f = 0; f + 1 < it.n_frames; f++) {
			const FramePlan &fp = b->frames[it.first_frame + f];
			if (!(fp.independent || fp.n_blocks <= 1) && fp.n_blocks) {
				b->presize.push_back(fp.first_block + fp.n_blocks - 1);
				b->presize_max.push_back(fp.block_max);
			}
		}
	if (!b->presize.empty()) b->placed = false;
	return LZ4ADA_OK;
}

const char *lz4ada_batch_k1_kernel_name(const lz4ada_batch *b) { return b ? b->k1_name : ""; }

uint32_t lz4ada_batch_retried_streams(const lz4ada_batch *b)
{
	uint32_t n = 0;
	if (b)
		for (const ItemPlan &it : b->items)
			if (it.slow) n++;
	return n;
}

int lz4ada_batch_kernel_ms(const lz4ada_batch *b, float ms[3])
{
	if (!b || !ms) return LZ4ADA_ASSERTION_ERROR;
	for (int k = 0; k < 3; k++) ms[k] = b->kernel_ms[k];
	return LZ4ADA_OK;
}

int lz4ada_batch_upload(lz4ada_batch *b, const uint8_t *src_host, uint8_t *src_dev)
{
	if (!b) return LZ4ADA_ASSERTION_ERROR;
	if (!b->ctx) {
		Raised why;
		b->ctx = default_context(&why);   // comes with the batch's reference
		if (!b->ctx) return LZ4ADA_DEVICE_ERROR;
	}
	lz4b200_ctx *ctx = b->ctx;
	const size_t nb = b->descs.size();
	// (only the bytes the batch's streams occupy: a share of a bigger buffer copies its share)
	if (src_host && src_dev && b->span_hi > b->span_lo &&
	    lz4b200_h2d(ctx, src_dev + b->span_lo, src_host + b->span_lo, b->span_hi - b->span_lo) != LZ4B200_OK)
		return LZ4ADA_DEVICE_ERROR;
	if (!b->d_desc && nb) {
		if (lz4b200_alloc(ctx, sizeof(lz4b200_blk_desc) * nb, reinterpret_cast<void **>(&b->d_desc)) != LZ4B200_OK ||
		    lz4b200_alloc(ctx, sizeof(lz4b200_blk_status) * nb, reinterpret_cast<void **>(&b->d_status)) != LZ4B200_OK ||
		    lz4b200_alloc_host(ctx, sizeof(lz4b200_blk_status) * nb, reinterpret_cast<void **>(&b->h_status)) != LZ4B200_OK)
			return LZ4ADA_DEVICE_ERROR;
	}
	if (!b->placed && nb) {
		// K5 over the blocks that decide frame bases (needs the compressed bytes on the device)
		if (!src_dev) return LZ4ADA_ASSERTION_ERROR;
		std::vector<lz4b200_blk_desc> pd(b->presize.size());
		for (size_t i = 0; i < pd.size(); i++) pd[i] = b->descs[b->presize[i]];
		lz4b200_blk_desc *d_pd = nullptr;
		lz4b200_blk_status *d_ps = nullptr;
		std::vector<lz4b200_blk_status> ps(pd.size());
		if (lz4b200_alloc(ctx, sizeof(lz4b200_blk_desc) * pd.size(), reinterpret_cast<void **>(&d_pd)) != LZ4B200_OK ||
		    lz4b200_alloc(ctx, sizeof(lz4b200_blk_status) * pd.size(), reinterpret_cast<void **>(&d_ps)) != LZ4B200_OK)
			return LZ4ADA_DEVICE_ERROR;
		const bool bad = lz4b200_h2d(ctx, d_pd, pd.data(), sizeof(lz4b200_blk_desc) * pd.size()) != LZ4B200_OK ||
				 lz4b200_size_blocks(ctx, src_dev, uint32_t(pd.size()), d_pd, d_ps) != LZ4B200_OK ||
				 lz4b200_d2h(ctx, ps.data(), d_ps, sizeof(lz4b200_blk_status) * pd.size()) != LZ4B200_OK ||
				 lz4b200_sync(ctx) != LZ4B200_OK;
		lz4b200_free(ctx, d_pd);
		lz4b200_free(ctx, d_ps);
		if (bad) return LZ4ADA_DEVICE_ERROR;
		std::vector<uint32_t> &sized = b->sized;
		sized.assign(pd.size(), 0);
		b->have_sized = true;
		for (size_t i = 0; i < pd.size(); i++) {
			const uint32_t bm = b->presize_max[i];   // the owning frame's block maximum, recorded at plan time
			// a block K5 cannot walk to its end keeps its block-maximum slot: K1 reports what is wrong with it
			sized[i] = ps[i].code == LZ4B200_ST_OK ? std::min(ps[i].out_len, bm) : (b->exact_sizing ? 0xffffffffu : std::min(ps[i].out_len, bm));
		}
		const uint64_t upper = b->out_bytes;
		place(b, &sized);
		b->out_bytes = std::max(b->out_bytes, upper);   // never shrink what the caller allocated from
		b->tables_uploaded = false;
	}
	if (!b->tables_uploaded && nb) {
		const size_t nh = b->hash_frames.size(), nc = b->chains.size();
		if (nc > b->cap_chains) {
			if (b->d_chains) lz4b200_free(ctx, b->d_chains);
			b->cap_chains = nc;
			if (lz4b200_alloc(ctx, sizeof(lz4b200_chain) * nc, reinterpret_cast<void **>(&b->d_chains)) != LZ4B200_OK)
				return LZ4ADA_DEVICE_ERROR;
		}
		if (nh && !b->d_hash_frames) {
			if (lz4b200_alloc(ctx, sizeof(lz4b200_frame_blocks) * nh, reinterpret_cast<void **>(&b->d_hash_frames)) != LZ4B200_OK ||
			    lz4b200_alloc(ctx, 8 * nh, reinterpret_cast<void **>(&b->d_digest)) != LZ4B200_OK ||
			    lz4b200_alloc_host(ctx, 8 * nh, reinterpret_cast<void **>(&b->h_digest)) != LZ4B200_OK)
				return LZ4ADA_DEVICE_ERROR;
		}
		const size_t n2 = b->k2_idx.size();
		if (n2 > b->cap_k2) {
			if (b->d_k2_idx) { lz4b200_free(ctx, b->d_k2_idx); lz4b200_free(ctx, b->d_k2_spans); lz4b200_free(ctx, b->d_k2_scratch); }
			b->d_k2_idx = nullptr; b->d_k2_spans = nullptr; b->d_k2_scratch = nullptr;
			b->cap_k2 = 0;
			if (lz4b200_alloc(ctx, 4 * n2, reinterpret_cast<void **>(&b->d_k2_idx)) != LZ4B200_OK ||
			    lz4b200_alloc(ctx, sizeof(lz4b200_hash_span) * n2, reinterpret_cast<void **>(&b->d_k2_spans)) != LZ4B200_OK ||
			    lz4b200_alloc(ctx, 4 * n2, reinterpret_cast<void **>(&b->d_k2_scratch)) != LZ4B200_OK)
				return LZ4ADA_DEVICE_ERROR;
			b->cap_k2 = n2;
		}
		if (n2 && (lz4b200_h2d(ctx, b->d_k2_idx, b->k2_idx.data(), 4 * n2) != LZ4B200_OK ||
			   lz4b200_h2d(ctx, b->d_k2_spans, b->k2_spans.data(), sizeof(lz4b200_hash_span) * n2) != LZ4B200_OK))
			return LZ4ADA_DEVICE_ERROR;
		if (lz4b200_h2d(ctx, b->d_desc, b->descs.data(), sizeof(lz4b200_blk_desc) * nb) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		if (nc && lz4b200_h2d(ctx, b->d_chains, b->chains.data(), sizeof(lz4b200_chain) * nc) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		if (nh && lz4b200_h2d(ctx, b->d_hash_frames, b->hash_frames.data(), sizeof(lz4b200_frame_blocks) * nh) != LZ4B200_OK)
			return LZ4ADA_DEVICE_ERROR;
		if (lz4b200_sync(ctx) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		b->tables_uploaded = true;
	}
	return LZ4ADA_OK;
}

int lz4ada_batch_run(lz4ada_batch *b, const uint8_t *src_dev, uint8_t *dst_dev)
{
	if (!b || !b->ctx) return LZ4ADA_ASSERTION_ERROR;
	lz4b200_ctx *ctx = b->ctx;
	const size_t nb = b->descs.size(), nh = b->hash_frames.size(), nc = b->chains.size();
	if (nb) {
		if (!b->tables_uploaded) return LZ4ADA_ASSERTION_ERROR;
		bool any_slow_before = false;
		for (ItemPlan &it : b->items)
			if (it.slow) any_slow_before = true;
		if (any_slow_before) {
			// a previous run re-placed some streams; restore the optimistic tables
			place(b, b->have_sized ? &b->sized : nullptr);
			if (lz4b200_h2d(ctx, b->d_desc, b->descs.data(), sizeof(lz4b200_blk_desc) * nb) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		}
		for (void *&e : b->ev)
			if (!e && lz4b200_event_create(ctx, &e) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		// every block's status starts out as "not run" (code 0xffffffff): a block no kernel of this run reached can
		// never be taken for decoded on the strength of an earlier run's status
		if (lz4b200_memset(ctx, b->d_status, 0xff, sizeof(lz4b200_blk_status) * nb) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		lz4b200_event_record(ctx, b->ev[0]);
		{
			// K1's own rule picks the lane-per-block kernel (v6) from the block count alone.  What its lanes are good at,
			// though, are blocks of up to 256 KiB that take long -- compressed ones of some size; stored blocks and
			// RLE-like ones are over in microseconds (and in v6 park a whole warp meanwhile), and a multi-megabyte block
			// in ONE lane would take a second.  With too few suitable blocks the launch lasts as long as one of them
			// while most lanes idle (2 GiB of thirds text / RLE / random in 64 KiB blocks: 17.5 ms lane-per-block,
			// 5.6 ms warp-per-block), so the host, which has the table, decides.
			if (b->heavy_blocks < 0) b->heavy_blocks = int64_t(heavy_blocks(b, 0, nb));
			const int rc1 = launch_k1(b, src_dev, dst_dev, 0, nb, uint64_t(b->heavy_blocks));
			if (rc1 != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
			if (launch_k2(b, src_dev, dst_dev, 0, nb) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		}
		lz4b200_event_record(ctx, b->ev[1]);
		if (nc && lz4b200_decode_linked(ctx, src_dev, dst_dev, uint32_t(nc), b->d_chains, b->d_desc, b->d_status) != LZ4B200_OK)
			return LZ4ADA_DEVICE_ERROR;
		lz4b200_event_record(ctx, b->ev[2]);
		if (nh && lz4b200_xxh32_frames(ctx, dst_dev, uint32_t(nh), b->d_hash_frames, b->d_desc, b->d_status, b->d_digest,
						b->d_digest + nh) != LZ4B200_OK)
			return LZ4ADA_DEVICE_ERROR;
		lz4b200_event_record(ctx, b->ev[3]);
		if (lz4b200_d2h(ctx, b->h_status, b->d_status, sizeof(lz4b200_blk_status) * nb) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		if (nh && lz4b200_d2h(ctx, b->h_digest, b->d_digest, 8 * nh) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		if (lz4b200_sync(ctx) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
		for (int k = 0; k < 3; k++) lz4b200_event_elapsed(ctx, b->ev[k], b->ev[k + 1], &b->kernel_ms[k]);
		if (!nc) b->kernel_ms[1] = 0;
		if (!nh) b->kernel_ms[2] = 0;
	}
	bool any_slow = false;
	for (ItemPlan &it : b->items) {
		it.slow = fold_item(b, it, false, [&](const FramePlan &fp, uint32_t &value) {
			if (fp.hash_slot == 0xffffffffu || !b->h_digest[nh + fp.hash_slot]) return false;
			value = b->h_digest[fp.hash_slot];
			return true;
		});
		if (it.slow) any_slow = true;
	}
	if (any_slow) {
		Raised r = run_slow_items(b, src_dev, dst_dev);
		if (r) return LZ4ADA_DEVICE_ERROR;
	}
	// traffic of this run: C + D + D for checksummed frames
	b->t_comp = b->t_out = b->t_reread = 0;
	for (const FramePlan &fp : b->frames) {
		uint64_t out = 0;
		for (uint32_t i = 0; i < fp.n_blocks; i++) {
			const lz4b200_blk_desc &d = b->descs[fp.first_block + i];
			b->t_comp += uint64_t(d.src_len) + 4 + ((d.flags & LZ4B200_BLK_HAS_CHECKSUM) ? 4 : 0);
			if (b->h_status && b->h_status[fp.first_block + i].code == LZ4B200_ST_OK) out += b->h_status[fp.first_block + i].out_len;
		}
		b->t_out += out;
		if (fp.has_cchk && fp.cchk_seen) b->t_reread += out;
	}
	return LZ4ADA_OK;
}

// Pipelined device stage for host buffers.  Chunks are runs of consecutive streams; the tables are
// already on the device (lz4ada_batch_upload with src_host == NULL), so each chunk is
// H2D(compressed) -> K1 -> K4 -> K3 -> D2H(status, digests) -> D2H(output) on its own lane.
int lz4ada_batch_run_pipelined(lz4ada_batch *b, const uint8_t *src_host, uint8_t *dst_host, uint8_t *src_dev,
			       uint8_t *dst_dev, uint32_t n_chunks)
{
	if (!b || !b->ctx || !src_host || !dst_host) return LZ4ADA_ASSERTION_ERROR;
	lz4b200_ctx *ctx = b->ctx;
	const size_t nb = b->descs.size(), nh = b->hash_frames.size(), nc = b->chains.size(), ni = b->items.size();
	if (!nb || !b->tables_uploaded || !b->placed) return LZ4ADA_ASSERTION_ERROR;
	if (n_chunks < 1) n_chunks = 1;
	if (n_chunks > ni) n_chunks = uint32_t(ni);
	// K1 picks its shape per launch from the chunk's block count (lz4b200_decode_blocks, tuning 0)
	const int saved_tuning = lz4b200_get_tuning(ctx);
	int rc = LZ4ADA_OK;
	size_t chain_pos = 0, hash_pos = 0;
	for (uint32_t c = 0; c < n_chunks && rc == LZ4ADA_OK; c++) {
		const size_t i0 = ni * c / n_chunks, i1 = ni * (c + 1) / n_chunks;
		if (i0 == i1) continue;
		const ItemPlan &first = b->items[i0], &last = b->items[i1 - 1];
		const uint32_t b0 = first.first_block, b1 = last.first_block + last.n_blocks;
		// source bytes of the chunk: streams are taken in the order given, usually back to back
		uint64_t s_lo = ~0ull, s_hi = 0, d_lo = ~0ull, d_hi = 0;
		for (size_t i = i0; i < i1; i++) {
			const ItemPlan &it = b->items[i];
			s_lo = std::min(s_lo, it.src_off);
			s_hi = std::max(s_hi, it.src_off + it.src_len);
			if (it.n_blocks) {
				d_lo = std::min(d_lo, it.dst_off);
				d_hi = std::max(d_hi, it.dst_off + it.dst_cap);
			}
		}
		size_t c0 = chain_pos, h0 = hash_pos;
		while (chain_pos < nc && b->chains[chain_pos].first_block < b1) chain_pos++;
		while (hash_pos < nh && b->hash_frames[hash_pos].first_block < b1) hash_pos++;
		const size_t ncc = chain_pos - c0, nhc = hash_pos - h0;
		bool bad = lz4b200_use_lane(ctx, int(c % 3) + 1) != LZ4B200_OK;
		bad = bad || lz4b200_h2d(ctx, src_dev + s_lo, src_host + s_lo, s_hi - s_lo) != LZ4B200_OK;
		if (b1 > b0)
			bad = bad || lz4b200_memset(ctx, b->d_status + b0, 0xff, sizeof(lz4b200_blk_status) * (b1 - b0)) != LZ4B200_OK;
		if (b1 > b0)
			bad = bad || launch_k1(b, src_dev, dst_dev, b0, b1, heavy_blocks(b, b0, b1)) != LZ4B200_OK ||
			      launch_k2(b, src_dev, dst_dev, b0, b1) != LZ4B200_OK;
		// chains and frame tables index blocks globally, so they get the un-offset arrays
		if (ncc)
			bad = bad || lz4b200_decode_linked(ctx, src_dev, dst_dev, uint32_t(ncc), b->d_chains + c0, b->d_desc, b->d_status) != LZ4B200_OK;
		if (nhc)
			bad = bad || lz4b200_xxh32_frames(ctx, dst_dev, uint32_t(nhc), b->d_hash_frames + h0, b->d_desc, b->d_status,
							   b->d_digest + h0, b->d_digest + nh + h0) != LZ4B200_OK;
		if (b1 > b0)
			bad = bad || lz4b200_d2h(ctx, b->h_status + b0, b->d_status + b0, sizeof(lz4b200_blk_status) * (b1 - b0)) != LZ4B200_OK;
		if (nhc) {
			bad = bad || lz4b200_d2h(ctx, b->h_digest + h0, b->d_digest + h0, 4 * nhc) != LZ4B200_OK;
			bad = bad || lz4b200_d2h(ctx, b->h_digest + nh + h0, b->d_digest + nh + h0, 4 * nhc) != LZ4B200_OK;
		}
		if (d_hi > d_lo) bad = bad || lz4b200_d2h(ctx, dst_host + d_lo, dst_dev + d_lo, d_hi - d_lo) != LZ4B200_OK;
		if (bad) rc = LZ4ADA_DEVICE_ERROR;
	}
	if (lz4b200_sync_all(ctx) != LZ4B200_OK) rc = LZ4ADA_DEVICE_ERROR;
	lz4b200_use_lane(ctx, 0);
	lz4b200_set_tuning(ctx, saved_tuning);
	if (rc != LZ4ADA_OK) return rc;
	bool any_slow = false;
	for (ItemPlan &it : b->items) {
		it.slow = fold_item(b, it, false, [&](const FramePlan &fp, uint32_t &value) {
			if (fp.hash_slot == 0xffffffffu || !b->h_digest[nh + fp.hash_slot]) return false;
			value = b->h_digest[fp.hash_slot];
			return true;
		});
		if (it.slow) any_slow = true;
	}
	if (any_slow) {
		if (run_slow_items(b, src_dev, dst_dev)) return LZ4ADA_DEVICE_ERROR;
		for (const ItemPlan &it : b->items)
			if (it.slow && it.out_len && lz4b200_d2h(ctx, dst_host + it.dst_off, dst_dev + it.dst_off, it.out_len) != LZ4B200_OK)
				return LZ4ADA_DEVICE_ERROR;
		if (lz4b200_sync(ctx) != LZ4B200_OK) return LZ4ADA_DEVICE_ERROR;
	}
	// The chunk copies brought back each stream's whole region (its size is only known after the fold): what lies
	// behind the bytes a stream produced is scratch of earlier calls, not output -- hand back zeros instead.
	for (const ItemPlan &it : b->items)
		if (it.n_blocks && it.dst_cap > it.out_len) memset(dst_host + it.dst_off + it.out_len, 0, it.dst_cap - it.out_len);
	return LZ4ADA_OK;
}

int lz4ada_batch_results(const lz4ada_batch *b, lz4ada_batch_result *results)
{
	if (!b || !results) return LZ4ADA_ASSERTION_ERROR;
	for (size_t k = 0; k < b->items.size(); k++) {
		const ItemPlan &it = b->items[k];
		results[k].exception = it.error.kind;
		results[k].end_of_frame = it.eof;
		results[k].n_frames = it.frames_seen;
		results[k].n_blocks = it.n_blocks;
		results[k].dst_off = it.dst_off;
		results[k].out_len = it.out_len;
	}
	return LZ4ADA_OK;
}

const char *lz4ada_batch_message(const lz4ada_batch *b, uint32_t item)
{
	if (!b || item >= b->items.size()) return "";
	return b->items[item].error.text.c_str();
}

void lz4ada_batch_free(lz4ada_batch *b) { delete b; }

// Device scratch of lz4ada_batch_decompress, kept between calls (cudaMalloc of several GB per call
// would dominate).  The pool belongs to its context (attached to it, shim.cu) and is freed when the
// context is torn down; the mutex serialises callers that share a context.
struct ScratchPool {
	lz4b200_ctx *ctx = nullptr;
	std::mutex busy;
	uint8_t *d_src = nullptr, *d_dst = nullptr;
	uint64_t cap_src = 0, cap_dst = 0;
	// table buffers handed to (and taken back from) the batch of the current call
	lz4b200_blk_desc *d_desc = nullptr;
	lz4b200_blk_status *d_status = nullptr, *h_status = nullptr;
	size_t cap_blocks = 0;
	lz4b200_frame_blocks *d_hash_frames = nullptr;
	uint32_t *d_digest = nullptr, *h_digest = nullptr;
	size_t cap_hash = 0;
	const char *last_k1 = "";   // K1 kernel of the last call's chunks
};

static void *pool_make(lz4b200_ctx *ctx)
{
	ScratchPool *p = new (std::nothrow) ScratchPool();
	if (p) p->ctx = ctx;   // no reference: the context owns the pool, not the other way round
	return p;
}

static void pool_free(lz4b200_ctx *ctx, void *obj)
{
	ScratchPool *p = static_cast<ScratchPool *>(obj);
	if (p->d_src) lz4b200_free(ctx, p->d_src);
	if (p->d_dst) lz4b200_free(ctx, p->d_dst);
	if (p->d_desc) lz4b200_free(ctx, p->d_desc);
	if (p->d_status) lz4b200_free(ctx, p->d_status);
	if (p->h_status) lz4b200_free_host(ctx, p->h_status);
	if (p->d_hash_frames) lz4b200_free(ctx, p->d_hash_frames);
	if (p->d_digest) lz4b200_free(ctx, p->d_digest);
	if (p->h_digest) lz4b200_free_host(ctx, p->h_digest);
	delete p;
}

static ScratchPool *pool_for(lz4b200_ctx *ctx) { return static_cast<ScratchPool *>(ctx_attachment(ctx, pool_make, pool_free)); }

static bool pool_reserve(ScratchPool *p, uint64_t src_bytes, uint64_t dst_bytes)
{
	if (src_bytes > p->cap_src) {
		if (p->d_src) lz4b200_free(p->ctx, p->d_src);
		p->d_src = nullptr;
		p->cap_src = 0;
		if (lz4b200_alloc(p->ctx, src_bytes + 64, reinterpret_cast<void **>(&p->d_src)) != LZ4B200_OK) return false;
		p->cap_src = src_bytes;
	}
	if (dst_bytes > p->cap_dst) {
		if (p->d_dst) lz4b200_free(p->ctx, p->d_dst);
		p->d_dst = nullptr;
		p->cap_dst = 0;
		if (lz4b200_alloc(p->ctx, dst_bytes + 64, reinterpret_cast<void **>(&p->d_dst)) != LZ4B200_OK) return false;
		p->cap_dst = dst_bytes;
	}
	return true;
}

int lz4ada_batch_decompress(lz4b200_ctx *ctx, const uint8_t *src_host, uint64_t src_bytes, uint8_t *dst_host,
			    uint64_t dst_bytes, uint32_t n_items, lz4ada_batch_item *items, int reservation,
			    lz4ada_batch_result *results, char *messages, size_t message_stride)
{
	lz4ada_batch *b = nullptr;
	int rc = lz4ada_batch_plan(ctx, src_host, src_bytes, n_items, items, reservation, &b);
	if (rc != LZ4ADA_OK) return rc;
	std::unique_ptr<lz4ada_batch> guard(b);
	if (!b->ctx) {
		Raised why;
		b->ctx = default_context(&why);   // comes with the batch's reference
		if (!b->ctx) return LZ4ADA_DEVICE_ERROR;
	}
	ctx = b->ctx;
	ScratchPool *pool = pool_for(ctx);
	if (!pool) return LZ4ADA_DEVICE_ERROR;
	std::lock_guard<std::mutex> pool_lock(pool->busy);
	const uint64_t need = lz4ada_batch_output_bytes(b);
	if (need > dst_bytes) return LZ4ADA_ASSERTION_ERROR;
	// spare room of the caller's buffer (up to 32 MiB of it) serves streams that outgrow their region
	const uint64_t capacity = need + std::min<uint64_t>(dst_bytes - need, 32ull << 20);
	b->dst_capacity = capacity;
	if (!pool_reserve(pool, b->span_hi - b->span_lo, capacity)) return LZ4ADA_DEVICE_ERROR;
	// device scratch holds [span_lo, span_hi) of the source: d_src is where byte 0 of the buffer would be
	uint8_t *d_src = pool->d_src - b->span_lo, *d_dst = pool->d_dst;
	// lend pooled table buffers to the batch when they are large enough (cudaMalloc / cudaMallocHost
	// per call cost ~100 ms, several times the device stage)
	const size_t nbk = b->descs.size(), nhk = b->hash_frames.size();
	const bool lend_blocks = nbk && pool->cap_blocks >= nbk, lend_hash = nhk && pool->cap_hash >= nhk;
	if (lend_blocks) { b->d_desc = pool->d_desc; b->d_status = pool->d_status; b->h_status = pool->h_status; }
	if (lend_hash) { b->d_hash_frames = pool->d_hash_frames; b->d_digest = pool->d_digest; b->h_digest = pool->h_digest; }
	struct Return {
		lz4ada_batch *b; ScratchPool *p; size_t nbk, nhk;
		~Return()
		{
			// keep the (possibly freshly allocated) buffers for the next call instead of freeing them
			if (b->d_desc && nbk >= p->cap_blocks) {
				if (p->d_desc && p->d_desc != b->d_desc) { lz4b200_free(p->ctx, p->d_desc); lz4b200_free(p->ctx, p->d_status); lz4b200_free_host(p->ctx, p->h_status); }
				p->d_desc = b->d_desc; p->d_status = b->d_status; p->h_status = b->h_status; p->cap_blocks = nbk;
			}
			if (b->d_desc == p->d_desc) { b->d_desc = nullptr; b->d_status = nullptr; b->h_status = nullptr; }
			if (b->d_hash_frames && nhk >= p->cap_hash) {
				if (p->d_hash_frames && p->d_hash_frames != b->d_hash_frames) { lz4b200_free(p->ctx, p->d_hash_frames); lz4b200_free(p->ctx, p->d_digest); lz4b200_free_host(p->ctx, p->h_digest); }
				p->d_hash_frames = b->d_hash_frames; p->d_digest = b->d_digest; p->h_digest = b->h_digest; p->cap_hash = nhk;
			}
			if (b->d_hash_frames == p->d_hash_frames) { b->d_hash_frames = nullptr; b->d_digest = nullptr; b->h_digest = nullptr; }
		}
	} give_back{b, pool, nbk, nhk};
	const uint64_t plain_guess = need;
	if (b->placed && b->descs.size()) {
		// fast shape: placement is known without touching the device -> pipeline H2D / kernels / D2H
		rc = lz4ada_batch_upload(b, nullptr, nullptr);   // tables only
		// ~512 MiB of output per chunk, at most 8 (LZ4ADA_E2E_CHUNKS / LZ4ADA_E2E_CHUNK_SHIFT: experiments)
		static const uint32_t max_chunks = [] { const char *e = getenv("LZ4ADA_E2E_CHUNKS"); return e ? uint32_t(atoi(e)) : 8u; }();
		static const uint32_t chunk_shift = [] { const char *e = getenv("LZ4ADA_E2E_CHUNK_SHIFT"); return e ? uint32_t(atoi(e)) : 29u; }();
		const uint32_t chunks = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(max_chunks, plain_guess >> chunk_shift)));
		if (rc == LZ4ADA_OK) rc = lz4ada_batch_run_pipelined(b, src_host, dst_host, d_src, d_dst, chunks);
	} else {
		rc = lz4ada_batch_upload(b, src_host, d_src);
		if (rc == LZ4ADA_OK) rc = lz4ada_batch_run(b, d_src, d_dst);
		if (rc == LZ4ADA_OK) {
			// bring back exactly what each stream produced
			for (const ItemPlan &it : b->items)
				if (it.out_len && lz4b200_d2h(ctx, dst_host + it.dst_off, d_dst + it.dst_off, it.out_len) != LZ4B200_OK) rc = LZ4ADA_DEVICE_ERROR;
			if (lz4b200_sync(ctx) != LZ4B200_OK) rc = LZ4ADA_DEVICE_ERROR;
		}
	}
	pool->last_k1 = b->k1_name;
	if (rc == LZ4ADA_OK && results) lz4ada_batch_results(b, results);
	if (rc == LZ4ADA_OK && messages && message_stride)
		for (uint32_t k = 0; k < n_items; k++) {
			const std::string &t = b->items[k].error.text;
			const size_t n = std::min(t.size(), message_stride - 1);
			memcpy(messages + size_t(k) * message_stride, t.data(), n);
			messages[size_t(k) * message_stride + n] = 0;
		}
	if (rc == LZ4ADA_OK)
		for (uint32_t k = 0; k < n_items; k++) {
			items[k].dst_off = b->items[k].dst_off;
			items[k].dst_cap = b->items[k].dst_cap;
		}
	return rc;
}

const char *lz4ada_last_k1_kernel_name(lz4b200_ctx *ctx)
{
	ScratchPool *p = ctx ? pool_for(ctx) : nullptr;
	return p ? p->last_k1 : "";
}

int lz4ada_batch_decompress_multi(uint32_t n_ctx, lz4b200_ctx *const *ctxs, const uint8_t *src_host, uint64_t src_bytes,
				  uint8_t *dst_host, uint64_t dst_bytes, uint32_t n_items, lz4ada_batch_item *items, int reservation,
				  lz4ada_batch_result *results, char *messages, size_t message_stride)
{
	if (n_ctx == 0 || !ctxs || (!items && n_items)) return LZ4ADA_ASSERTION_ERROR;
	for (uint32_t g = 0; g < n_ctx; g++)
		if (!ctxs[g]) return LZ4ADA_ASSERTION_ERROR;
	if (n_ctx == 1)
		return lz4ada_batch_decompress(ctxs[0], src_host, src_bytes, dst_host, dst_bytes, n_items, items, reservation, results, messages,
					       message_stride);
	// ---- deal whole streams to the devices: contiguous runs of the streams in the order given, cut where the running
	// sum of compressed bytes passes the next n-th of the total (a share's bytes then lie together in the source
	// buffer, so its H2D copies move its share and nothing else) ----
	std::vector<std::vector<uint32_t>> share(n_ctx);
	{
		uint64_t total = 0;
		for (uint32_t k = 0; k < n_items; k++) total += items[k].src_len + 1;
		uint64_t run = 0;
		for (uint32_t k = 0; k < n_items; k++) {
			const uint64_t mid = run + (items[k].src_len + 1) / 2;
			uint32_t g = uint32_t((unsigned __int128)mid * n_ctx / (total ? total : 1));
			if (g >= n_ctx) g = n_ctx - 1;
			share[g].push_back(k);
			run += items[k].src_len + 1;
		}
	}
	// ---- the output room of each device's share: a host-only plan per share (no device needed for that) ----
	std::vector<uint64_t> need(n_ctx, 0), base(n_ctx, 0);
	std::vector<std::vector<lz4ada_batch_item>> sub(n_ctx);
	std::vector<int> rcs(n_ctx, LZ4ADA_OK);
	auto plan_share = [&](uint32_t g) {
		g_plan_share = n_ctx;
		sub[g].resize(share[g].size());
		for (size_t j = 0; j < share[g].size(); j++) {
			sub[g][j] = items[share[g][j]];
			sub[g][j].dst_off = 0;
			sub[g][j].dst_cap = 0;   // planner-placed inside the share's region
		}
		lz4ada_batch *b = nullptr;
		rcs[g] = lz4ada_batch_plan(nullptr, src_host, src_bytes, uint32_t(sub[g].size()), sub[g].data(), reservation, &b);
		if (rcs[g] == LZ4ADA_OK) {
			need[g] = lz4ada_batch_output_bytes(b);
			lz4ada_batch_free(b);
		}
	};
	{
		std::vector<std::thread> th;
		for (uint32_t g = 1; g < n_ctx; g++) th.emplace_back(plan_share, g);
		plan_share(0);
		for (auto &t : th) t.join();
	}
	for (uint32_t g = 0; g < n_ctx; g++)
		if (rcs[g] != LZ4ADA_OK) return rcs[g];
	uint64_t cursor = 0;
	for (uint32_t g = 0; g < n_ctx; g++) {
		base[g] = align_up(cursor, 256);
		cursor = base[g] + need[g];
	}
	if (cursor > dst_bytes) return LZ4ADA_ASSERTION_ERROR;
	const uint64_t spare = (dst_bytes - cursor) / n_ctx;   // room for streams that outgrow their region, split evenly
	// ---- every device runs its share on a thread of its own ----
	std::vector<std::vector<lz4ada_batch_result>> res(n_ctx);
	std::vector<std::vector<char>> msg(n_ctx);
	auto run_share = [&](uint32_t g) {
		g_plan_share = n_ctx;
		res[g].resize(std::max<size_t>(1, share[g].size()));
		if (messages && message_stride) msg[g].assign(std::max<size_t>(1, share[g].size()) * message_stride, 0);
		if (share[g].empty()) return;
		// (the spare room of share g lies behind ALL planned regions, not behind its own: regions are packed)
		uint8_t *dst_g = dst_host + base[g];
		uint64_t cap_g = need[g];
		if (g + 1 == n_ctx) cap_g += spare * n_ctx;   // only the last region can grow in place
		rcs[g] = lz4ada_batch_decompress(ctxs[g], src_host, src_bytes, dst_g, cap_g, uint32_t(sub[g].size()), sub[g].data(), reservation,
						 res[g].data(), msg[g].empty() ? nullptr : msg[g].data(), message_stride);
	};
	{
		std::vector<std::thread> th;
		for (uint32_t g = 1; g < n_ctx; g++) th.emplace_back(run_share, g);
		run_share(0);
		for (auto &t : th) t.join();
	}
	for (uint32_t g = 0; g < n_ctx; g++)
		if (rcs[g] != LZ4ADA_OK) return rcs[g];
	for (uint32_t g = 0; g < n_ctx; g++)
		for (size_t j = 0; j < share[g].size(); j++) {
			const uint32_t k = share[g][j];
			if (results) {
				results[k] = res[g][j];
				results[k].dst_off += base[g];
			}
			items[k].dst_off = sub[g][j].dst_off + base[g];
			items[k].dst_cap = sub[g][j].dst_cap;
			if (messages && message_stride) memcpy(messages + size_t(k) * message_stride, msg[g].data() + j * message_stride, message_stride);
		}
	return LZ4ADA_OK;
}

}  // extern "C"
