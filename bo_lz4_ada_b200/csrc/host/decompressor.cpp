// decompressor.cpp -- the LZ4Ada package API over the device shim: Init, Init_With_Header,
// Init_For_Block, Update, Is_End_Of_Frame, To_Hex, XXHash32 (lib/lz4ada.ads:50-344).
//
// Update keeps the reference's contract: the caller's Buffer receives one block per call at the
// same ring position the reference would use (Output_First / Output_Last), and the same
// exception fires at the same byte.  What differs is where the work happens: a complete block
// goes H2D, is decoded by one warp against a device-resident 64 KiB history window, and the
// produced bytes come back D2H (lz4b200_stream_block).  There is no host-side LZ4 decoder.
//
// Read-ahead: when a block arrives straight out of a large Input and the
// independent blocks behind it are complete in that Input too, they are all decoded by ONE K1
// launch into a staging buffer; the following Update calls then only check that the caller presents
// the same bytes again, copy their block out of the staging buffer and let the stream adopt it
// (history window + running content checksum).  Every call still reports exactly one block, the same
// Num_Consumed / Output_First / Output_Last as the reference, and an erroneous block is never
// cached: it goes through the one-block path when its turn comes, with the same exception.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.hpp"

namespace lz4ada {

// ---- process-wide default device context -------------------------------------------------
// The slot holds one reference of its own on whatever context it names (lz4b200_retain), so a caller who
// destroys the context it registered -- or replaces the default while decompressors are alive -- leaves
// nothing dangling: every user of the slot leaves with a reference too.
static std::mutex g_ctx_mutex;
static lz4b200_ctx *g_ctx = nullptr;

lz4b200_ctx *default_context(Raised *why)
{
	std::lock_guard<std::mutex> lock(g_ctx_mutex);
	if (!g_ctx) {
		lz4b200_ctx *c = nullptr;
		if (lz4b200_create(0, nullptr, &c) != LZ4B200_OK) {
			if (why)
				*why = err_device("no CUDA device / context available -- this library has no CPU "
						  "decode path");
			return nullptr;
		}
		g_ctx = c;   // the creator's reference becomes the slot's
	}
	lz4b200_retain(g_ctx);
	return g_ctx;
}

// The read-ahead buffers of a streaming engine (device and pinned staging, tables, events).  A decompressor that is
// freed leaves its set with the device context, the next one made on that context takes it over: pinned allocations
// are what a new decompressor costs (~80 ms for the 64 MiB read-ahead), and a program that decompresses file after
// file makes one per file.  Owned by the context (second attachment), freed with it.
struct StageSet {
	uint8_t *d_src = nullptr, *h_copy = nullptr;
	uint8_t *d_stage[2] = {nullptr, nullptr}, *h_stage[2] = {nullptr, nullptr};
	lz4b200_blk_desc *d_desc = nullptr;
	lz4b200_blk_status *d_stat = nullptr;
	lz4b200_chain *d_chain = nullptr;
	void *hash_done[2] = {nullptr, nullptr};
	size_t cap_src = 0, cap_stage = 0, cap_n = 0;
};
static void free_stage_set(lz4b200_ctx *ctx, StageSet &b)
{
	if (b.d_src) lz4b200_free(ctx, b.d_src);
	if (b.h_copy) lz4b200_free_host(ctx, b.h_copy);
	for (int k = 0; k < 2; k++) {
		if (b.d_stage[k]) lz4b200_free(ctx, b.d_stage[k]);
		if (b.h_stage[k]) lz4b200_free_host(ctx, b.h_stage[k]);
		if (b.hash_done[k]) lz4b200_event_destroy(ctx, b.hash_done[k]);
	}
	if (b.d_desc) lz4b200_free(ctx, b.d_desc);
	if (b.d_stat) lz4b200_free(ctx, b.d_stat);
	if (b.d_chain) lz4b200_free(ctx, b.d_chain);
	b = StageSet();
}
struct EnginePool {
	std::mutex m;
	bool has = false;
	StageSet set;
};
static void *engine_pool_make(lz4b200_ctx *) { return new EnginePool(); }
static void engine_pool_free(lz4b200_ctx *ctx, void *p)
{
	EnginePool *e = static_cast<EnginePool *>(p);
	if (e->has) free_stage_set(ctx, e->set);
	delete e;
}

// ---- streaming engine: one block at a time on the device ---------------------------------
class DeviceStreamEngine : public BlockEngine {
public:
	explicit DeviceStreamEngine(int stream_block_max) : max_block_(uint32_t(stream_block_max)) {}
	~DeviceStreamEngine() override
	{
		if (ctx_) {
			join_hash_lane();   // nothing reads the staging buffers any more
			StageSet mine;
			mine.d_src = d_src_; mine.h_copy = h_copy_;
			for (int k = 0; k < 2; k++) { mine.d_stage[k] = d_stage_[k]; mine.h_stage[k] = h_stage_[k]; mine.hash_done[k] = hash_done_[k]; }
			mine.d_desc = d_desc_; mine.d_stat = d_stat_; mine.d_chain = d_chain_;
			mine.cap_src = cap_src_; mine.cap_stage = cap_stage_; mine.cap_n = cap_n_;
			EnginePool *pool = static_cast<EnginePool *>(ctx_attachment2(ctx_, engine_pool_make, engine_pool_free));
			bool kept = false;
			if (pool && (mine.cap_src || mine.cap_stage || mine.cap_n)) {
				std::lock_guard<std::mutex> lock(pool->m);
				if (!pool->has) {
					pool->set = mine;
					pool->has = true;
					kept = true;
				}
			}
			if (!kept) free_stage_set(ctx_, mine);
		}
		if (getenv("LZ4ADA_UPDATE_DEBUG") && dbg_n_served_)
			fprintf(stderr, "[lz4ada update] %ld read-aheads %.1f ms (scan %.1f, reserve %.1f, enqueue %.1f, sync %.1f, keep %.1f), %ld blocks served %.1f ms; all of block() %.1f ms, content checksums %.1f ms\n",
				dbg_n_ahead_, dbg_ahead_ * 1e3, dbg_stage_[0] * 1e3, dbg_stage_[1] * 1e3, dbg_stage_[2] * 1e3, dbg_stage_[3] * 1e3, dbg_stage_[4] * 1e3,
				dbg_n_served_, dbg_serve_ * 1e3, dbg_stage_[5] * 1e3, dbg_digest_ * 1e3);
		if (stream_) lz4b200_stream_destroy(stream_);
		if (ctx_) lz4b200_destroy(ctx_);   // the engine's reference (default_context took it)
	}

	Raised new_frame(Walker &) override   // Reset_Outer_For_Next_Frame, lib/lz4ada.adb:451-461
	{
		output_pos_ = 0;
		drop_ahead();
		pend_off_.clear();   // (the old frame's checksum, if it had one, has been asked for by now: content_checksum flushes)
		pend_len_.clear();
		if (stream_ && !join_hash_lane()) return device_failure();
		if (stream_ && lz4b200_stream_reset(stream_) != LZ4B200_OK) return device_failure();
		return ok();
	}

	Raised block(Walker &w, const uint8_t *blk, int blk_len, uint8_t *buffer, int buffer_len, int &of,
		     int &ol) override   // Decode_Full_Block_With_Trailer, :661-696
	{
		const double t_in = dbg_now();
		const Raised r = block_inner(w, blk, blk_len, buffer, buffer_len, of, ol);
		dbg_stage_[5] += dbg_now() - t_in;
		return r;
	}

	Raised block_inner(Walker &w, const uint8_t *blk, int blk_len, uint8_t *buffer, int buffer_len, int &of, int &ol)
	{
		if (Raised r = ensure_stream()) return r;
		const int raw_len = blk_len - w.m.block_checksum_length;
		if (output_pos_ >= kHistorySize) output_pos_ = 0;   // ring cursor, :678-680
		int64_t cap = int64_t(buffer_len) - output_pos_;
		if (cap < 0) cap = 0;
		if (cap > int64_t(max_block_)) cap = max_block_;
		uint32_t flags = 0;
		if (!w.m.is_compressed) flags |= LZ4B200_BLK_STORED;
		if (w.m.block_checksum_length) flags |= LZ4B200_BLK_HAS_CHECKSUM;
		lz4b200_blk_status st;
		memset(&st, 0, sizeof st);
		// ---- read-ahead: served from the staging buffer, or decoded together with the blocks behind it ----
		bool served = false;
		if (ahead_pos_ < ahead_.size()) {
			const Ahead &e = ahead_[ahead_pos_];
			if (e.blk_len == blk_len && e.flags == flags && memcmp(blk, h_copy_ + e.copy_off, size_t(blk_len)) == 0)
				served = true;
			else
				drop_ahead();   // the caller came back with something else
		}
		if (!served && ahead_pos_ >= ahead_.size() && w.lookahead) {
			const double t0 = dbg_now();
			if (!flush_pending()) return device_failure();   // (the staging buffer is about to be overwritten)
			if (read_ahead(w, blk, blk_len, flags)) served = true;
			dbg_ahead_ += dbg_now() - t0;
			dbg_n_ahead_++;
		}
		if (served) {
			const double t0 = dbg_now();
			dbg_n_served_++;
			const Ahead &e = ahead_[ahead_pos_];
			if (int64_t(e.st.out_len) <= cap) {
				st = e.st;
				if (st.out_len) memcpy(buffer + output_pos_, h_stage_[cur_] + e.out_off, st.out_len);
				// the stream adopts the block (history window + running content checksum) later, together with the
				// other blocks served from this read-ahead: one hash launch and one window copy for all of them
				pend_off_.push_back(e.out_off);
				pend_len_.push_back(st.out_len);
				pend_hash_ = w.m.content_checksum_length != 0;
				pend_buf_ = cur_;
				ahead_pos_++;
				dbg_serve_ += dbg_now() - t0;
			} else {
				served = false;   // does not fit the caller's Buffer here: the one-block path reports it
				drop_ahead();
			}
		}
		const double t_single = dbg_now();
		if (!served && (!flush_pending() || !join_hash_lane())) return device_failure();
		if (!served) dbg_join_ += dbg_now() - t_single;
		if (!served &&
		    lz4b200_stream_block2(stream_, blk, uint32_t(raw_len), flags, w.m.content_checksum_length != 0,
					  buffer + output_pos_, uint32_t(cap), uint32_t(w.m.frame_block_max), &st) != LZ4B200_OK)
			return device_failure();
		if (!served) {
			dbg_single_ += dbg_now() - t_single;
			dbg_n_single_++;
		}
		// Decrease_Data_Size_Remaining (:826-839) fires inside Write_Output, i.e. before any
		// later check of the same block; the block checksum (:672-676) comes before everything.
		if (w.m.has_content_size && st.code != LZ4B200_ST_BLOCK_CHECKSUM) {
			const uint64_t produced = st.code == LZ4B200_ST_OK ? st.out_len : st.err_pos;
			if (w.m.size_remaining < produced) return err_content_size_exceeded();
		}
		if (st.code != LZ4B200_ST_OK) return status_to_raised(st, buffer_len);
		if (w.m.has_content_size) w.m.size_remaining -= st.out_len;
		of = output_pos_;
		ol = output_pos_ + int(st.out_len) - 1;
		output_pos_ += int(st.out_len);
		return ok();
	}

	Raised content_checksum(Walker &, uint32_t declared) override   // :493-511
	{
		if (Raised r = ensure_stream()) return r;
		const double t_in = dbg_now();
		if (!flush_pending() || !join_hash_lane()) return device_failure();
		uint32_t computed = 0;
		if (lz4b200_stream_digest(stream_, &computed) != LZ4B200_OK) return device_failure();
		dbg_digest_ += dbg_now() - t_in;
		if (getenv("LZ4ADA_UPDATE_DEBUG"))
			fprintf(stderr, "[lz4ada update, cumulative at frame end] read-aheads %ld: %.1f ms (pinned copy + buffers %.1f, enqueue %.1f, sync %.1f); served %ld: %.1f ms; one-block path %ld: %.1f ms (of which waiting for the hash lane %.1f); all of block() %.1f ms; content checksums %.1f ms\n",
				dbg_n_ahead_, dbg_ahead_ * 1e3, dbg_stage_[1] * 1e3, dbg_stage_[2] * 1e3, dbg_stage_[3] * 1e3, dbg_n_served_, dbg_serve_ * 1e3,
				dbg_n_single_, dbg_single_ * 1e3, dbg_join_ * 1e3, dbg_stage_[5] * 1e3, dbg_digest_ * 1e3);
		if (computed != declared) return err_content_checksum(computed, declared);
		return ok();
	}

	Raised frame_ended(Walker &w) override   // Set_Frame_Has_Ended, :465-477
	{
		if (w.m.has_content_size && w.m.size_remaining != 0) return err_content_size_left(w.m.size_remaining);
		return ok();
	}

private:
	// ---- read-ahead -----------------------------------------------------------------------
	struct Ahead {
		size_t copy_off;          // payload (+ trailer) of the block in h_copy_
		int blk_len;
		uint32_t flags;
		uint32_t out_off;         // its decoded bytes in the staging buffers
		lz4b200_blk_status st;    // always LZ4B200_ST_OK
	};
	static constexpr int kAheadMaxBlocks = 255;               // (what lz4b200_stream_adopt_list takes at once)
	static constexpr uint64_t kAheadMaxBytes = 64ull << 20;   // decoded bytes per read-ahead

	void drop_ahead()
	{
		ahead_.clear();
		ahead_pos_ = 0;
	}

	// (LZ4ADA_UPDATE_DEBUG=1: where the time of the read-ahead path goes, printed when the decompressor is freed)
	static double dbg_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	double dbg_ahead_ = 0, dbg_serve_ = 0, dbg_digest_ = 0, dbg_stage_[6] = {0, 0, 0, 0, 0, 0};
	long dbg_n_ahead_ = 0, dbg_n_served_ = 0, dbg_n_single_ = 0;
	double dbg_single_ = 0, dbg_join_ = 0;
	// Blocks served from the staging buffer that the stream has not adopted yet (offsets into d_stage_, in order).
	std::vector<uint32_t> pend_off_, pend_len_;
	bool pend_hash_ = false;
	int pend_buf_ = 0;
	bool hash_lane_busy_ = false;
	void *hash_done_[2] = {nullptr, nullptr};   // events: the hash lane has passed the adoption that reads staging buffer k
	bool hash_reads_[2] = {false, false};
	// The adoption (one hash launch over all pending blocks + the window copy) runs on a lane of its own, so that the
	// serial content checksum of one read-ahead (8 ms for 16 MiB) overlaps with the decode of the next; the two
	// staging buffers alternate, and whoever needs the stream's state on the main lane joins the hash lane first.
	bool flush_pending()
	{
		if (pend_off_.empty()) return true;
		bool fine = lz4b200_use_lane(ctx_, 1) == LZ4B200_OK;
		fine = fine && lz4b200_stream_adopt_list(stream_, d_stage_[pend_buf_], uint32_t(pend_off_.size()), pend_off_.data(), pend_len_.data(),
							   pend_hash_ ? 1 : 0) == LZ4B200_OK;
		if (fine && !hash_done_[pend_buf_]) fine = lz4b200_event_create(ctx_, &hash_done_[pend_buf_]) == LZ4B200_OK;
		fine = fine && lz4b200_event_record(ctx_, hash_done_[pend_buf_]) == LZ4B200_OK;
		lz4b200_use_lane(ctx_, 0);
		hash_lane_busy_ = true;
		hash_reads_[pend_buf_] = true;
		pend_off_.clear();
		pend_len_.clear();
		return fine;
	}
	bool join_hash_lane()
	{
		if (!hash_lane_busy_) return true;
		bool fine = lz4b200_use_lane(ctx_, 1) == LZ4B200_OK && lz4b200_sync(ctx_) == LZ4B200_OK;
		lz4b200_use_lane(ctx_, 0);
		hash_lane_busy_ = false;
		hash_reads_[0] = hash_reads_[1] = false;
		return fine;
	}

	// Decode `blk` and the complete independent blocks behind it in the caller's Input with one K1 launch.
	// true = the cache now starts with `blk` (decoded fine); false = nothing cached, use the one-block path.
	bool read_ahead(Walker &w, const uint8_t *blk, int blk_len, uint32_t flags)
	{
		drop_ahead();
		const double ta = dbg_now();
		if (w.m.format != Format::Modern || !w.m.block_independent || w.lookahead != blk + blk_len) return false;
		const int bc = w.m.block_checksum_length;
		const uint32_t block_max = uint32_t(w.m.frame_block_max);
		if (block_max == 0 || block_max > max_block_) return false;
		const uint32_t stride = (block_max + 255u) & ~255u;
		struct Item { size_t off; int len; uint32_t flags; };
		std::vector<Item> items;
		items.push_back({0, blk_len, flags});
		const uint8_t *p = w.lookahead, *end = w.lookahead + w.lookahead_len;
		while (int(items.size()) < kAheadMaxBlocks && uint64_t(items.size() + 1) * stride <= kAheadMaxBytes) {
			// the size word as Try_Detect_Input_Length reads it (lib/lz4ada.adb:525-585)
			if (end - p < kBlockSizeBytes) break;
			uint32_t word = load32(p);
			if (word == 0) break;   // end mark
			const bool stored = (word & 0x80000000u) != 0;
			word &= 0x7ffffffu;
			if (int64_t(word) + kBlockSizeBytes + bc > int64_t(w.input_buffer_len)) break;   // the walker raises this one
			const int total = int(word) + bc;
			if (end - (p + kBlockSizeBytes) < total) break;   // not complete in this Input
			uint32_t f = stored ? LZ4B200_BLK_STORED : 0u;
			if (bc) f |= LZ4B200_BLK_HAS_CHECKSUM;
			items.push_back({size_t(p + kBlockSizeBytes - blk), total, f});
			p += kBlockSizeBytes + total;
		}
		if (items.size() < 2) return false;
		const size_t n = items.size();
		const size_t span = items.back().off + size_t(items.back().len);
		const double tb = dbg_now();
		// the other staging buffer: the one just served from may still be read by the hash lane.  (The hash that read
		// THIS buffer was queued a whole read-ahead ago.)
		cur_ ^= 1;
		if (hash_reads_[cur_]) {
			if (lz4b200_event_sync(ctx_, hash_done_[cur_]) != LZ4B200_OK) return false;
			hash_reads_[cur_] = false;
		}
		if (!reserve(span + 64, n * size_t(stride) + 64, n)) return false;
		// the caller's bytes: a pinned copy serves the H2D (asynchronous for real) and the "same bytes again?" checks
		memcpy(h_copy_, blk, span);
		const double tc = dbg_now();
		std::vector<lz4b200_blk_desc> descs(n);
		for (size_t i = 0; i < n; i++) {
			descs[i].src_off = items[i].off;
			descs[i].src_len = uint32_t(items[i].len - bc);
			descs[i].dst_off = uint64_t(i) * stride;
			descs[i].dst_cap = block_max;
			descs[i].flags = items[i].flags;
			descs[i].hist_avail = 0;
		}
		std::vector<lz4b200_blk_status> stats(n);
		// big blocks (>= 64 KiB of compressed bytes: a 1 - 4 MiB block) go to the chain kernel as chains of one, a CTA
		// each (kernels_k7.cuh), instead of a warp each: 18 ms against 40 ms for a 4 MiB text block
		bool big = false;
		for (size_t i = 0; i < n; i++) big = big || (descs[i].src_len >= 65536u && !(descs[i].flags & LZ4B200_BLK_STORED));
		std::vector<lz4b200_chain> chains;
		if (big) {
			chains.resize(n);
			for (size_t i = 0; i < n; i++) {
				descs[i].flags |= LZ4B200_BLK_CHAINED | LZ4B200_BLK_FIRST_OF_FRAME | LZ4B200_BLK_SOLO;
				chains[i].first_block = uint32_t(i);
				chains[i].n_blocks = 1;
				chains[i].dst_off = descs[i].dst_off;
				chains[i].dst_cap = descs[i].dst_cap;
			}
			if (lz4b200_h2d(ctx_, d_chain_, chains.data(), sizeof(lz4b200_chain) * n) != LZ4B200_OK) return false;
		}
		if (lz4b200_h2d(ctx_, d_src_, h_copy_, span) != LZ4B200_OK ||
		    lz4b200_h2d(ctx_, d_desc_, descs.data(), sizeof(lz4b200_blk_desc) * n) != LZ4B200_OK ||
		    (big ? lz4b200_memset(ctx_, d_stat_, 0xff, sizeof(lz4b200_blk_status) * n) : LZ4B200_OK) != LZ4B200_OK ||
		    (big ? lz4b200_decode_linked(ctx_, d_src_, d_stage_[cur_], uint32_t(n), d_chain_, d_desc_, d_stat_)
			 : lz4b200_decode_blocks(ctx_, d_src_, d_stage_[cur_], uint32_t(n), d_desc_, d_stat_)) != LZ4B200_OK ||
		    lz4b200_d2h(ctx_, stats.data(), d_stat_, sizeof(lz4b200_blk_status) * n) != LZ4B200_OK ||
		    lz4b200_d2h(ctx_, h_stage_[cur_], d_stage_[cur_], n * size_t(stride)) != LZ4B200_OK)
			return false;
		const double td = dbg_now();
		if (lz4b200_sync(ctx_) != LZ4B200_OK) return false;
		const double te = dbg_now();
		// keep the leading run of good blocks; the first one that is not (an error, or a block that reaches into
		// its predecessor) and everything behind it take the one-block path when their turn comes
		size_t good = 0;
		while (good < n && stats[good].code == LZ4B200_ST_OK) good++;
		if (good == 0) return false;
		for (size_t i = 0; i < good; i++) ahead_.push_back({items[i].off, items[i].len, items[i].flags, uint32_t(i) * stride, stats[i]});
		const double tf = dbg_now();
		dbg_stage_[0] += tb - ta; dbg_stage_[1] += tc - tb; dbg_stage_[2] += td - tc; dbg_stage_[3] += te - td; dbg_stage_[4] += tf - te;
		return true;
	}

	bool reserve(size_t src_bytes, size_t stage_bytes, size_t n)
	{
		if (src_bytes > cap_src_) {
			if (d_src_) lz4b200_free(ctx_, d_src_);
			if (h_copy_) lz4b200_free_host(ctx_, h_copy_);
			d_src_ = h_copy_ = nullptr;
			cap_src_ = 0;
			if (lz4b200_alloc(ctx_, src_bytes, reinterpret_cast<void **>(&d_src_)) != LZ4B200_OK ||
			    lz4b200_alloc_host(ctx_, src_bytes, reinterpret_cast<void **>(&h_copy_)) != LZ4B200_OK)
				return false;
			cap_src_ = src_bytes;
		}
		if (stage_bytes > cap_stage_) {
			if (!join_hash_lane()) return false;   // (nobody reads what is about to be freed)
			for (int k = 0; k < 2; k++) {
				if (d_stage_[k]) lz4b200_free(ctx_, d_stage_[k]);
				if (h_stage_[k]) lz4b200_free_host(ctx_, h_stage_[k]);
				d_stage_[k] = h_stage_[k] = nullptr;
			}
			cap_stage_ = 0;
			for (int k = 0; k < 2; k++)
				if (lz4b200_alloc(ctx_, stage_bytes, reinterpret_cast<void **>(&d_stage_[k])) != LZ4B200_OK ||
				    lz4b200_alloc_host(ctx_, stage_bytes, reinterpret_cast<void **>(&h_stage_[k])) != LZ4B200_OK)
					return false;
			cap_stage_ = stage_bytes;
		}
		if (n > cap_n_) {
			if (d_desc_) lz4b200_free(ctx_, d_desc_);
			if (d_stat_) lz4b200_free(ctx_, d_stat_);
			if (d_chain_) lz4b200_free(ctx_, d_chain_);
			d_desc_ = nullptr;
			d_stat_ = nullptr;
			d_chain_ = nullptr;
			cap_n_ = 0;
			if (lz4b200_alloc(ctx_, sizeof(lz4b200_blk_desc) * n, reinterpret_cast<void **>(&d_desc_)) != LZ4B200_OK ||
			    lz4b200_alloc(ctx_, sizeof(lz4b200_blk_status) * n, reinterpret_cast<void **>(&d_stat_)) != LZ4B200_OK ||
			    lz4b200_alloc(ctx_, sizeof(lz4b200_chain) * n, reinterpret_cast<void **>(&d_chain_)) != LZ4B200_OK)
				return false;
			cap_n_ = n;
		}
		return true;
	}

	std::vector<Ahead> ahead_;
	size_t ahead_pos_ = 0;
	uint8_t *d_src_ = nullptr, *h_copy_ = nullptr;   // the Input bytes of the read-ahead, on the device and pinned on the host
	uint8_t *d_stage_[2] = {nullptr, nullptr}, *h_stage_[2] = {nullptr, nullptr};   // decoded blocks, two buffers taking turns
	int cur_ = 0;
	lz4b200_blk_desc *d_desc_ = nullptr;
	lz4b200_blk_status *d_stat_ = nullptr;
	lz4b200_chain *d_chain_ = nullptr;
	size_t cap_src_ = 0, cap_stage_ = 0, cap_n_ = 0;

	Raised ensure_stream()
	{
		if (stream_) return ok();
		Raised why;
		if (!ctx_) ctx_ = default_context(&why);
		if (!ctx_) return why;
		if (lz4b200_stream_create(ctx_, max_block_, &stream_) != LZ4B200_OK) {
			stream_ = nullptr;
			return device_failure();
		}
		if (EnginePool *pool = static_cast<EnginePool *>(ctx_attachment2(ctx_, engine_pool_make, engine_pool_free))) {
			std::lock_guard<std::mutex> lock(pool->m);
			if (pool->has) {   // the buffers a decompressor freed earlier left behind
				const StageSet &b = pool->set;
				d_src_ = b.d_src; h_copy_ = b.h_copy;
				for (int k = 0; k < 2; k++) { d_stage_[k] = b.d_stage[k]; h_stage_[k] = b.h_stage[k]; hash_done_[k] = b.hash_done[k]; }
				d_desc_ = b.d_desc; d_stat_ = b.d_stat; d_chain_ = b.d_chain;
				cap_src_ = b.cap_src; cap_stage_ = b.cap_stage; cap_n_ = b.cap_n;
				pool->set = StageSet();
				pool->has = false;
			}
		}
		return ok();
	}
	Raised device_failure() { return err_device(ctx_ ? lz4b200_last_error(ctx_) : "no device context"); }

	uint32_t max_block_;
	lz4b200_ctx *ctx_ = nullptr;
	lz4b200_stream *stream_ = nullptr;
	int output_pos_ = 0;   // Output_Pos of the reference's ring
};

}  // namespace lz4ada

using namespace lz4ada;

struct lz4ada_decompressor {
	DeviceStreamEngine engine;
	Walker walker;
	std::string message;
	lz4ada_decompressor(const Meta &m, int in_last, int min_buffer_size)
		: engine(min_buffer_size), walker(m, in_last, &engine)
	{
	}
};

static void copy_message(char *dst, size_t cap, const std::string &s)
{
	if (!dst || !cap) return;
	const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
	memcpy(dst, s.data(), n);
	dst[n] = 0;
}

extern "C" {

int lz4ada_set_device_context(lz4b200_ctx *ctx)
{
	std::lock_guard<std::mutex> lock(g_ctx_mutex);
	if (g_ctx == ctx) return LZ4ADA_OK;
	if (ctx) lz4b200_retain(ctx);
	if (g_ctx) lz4b200_destroy(g_ctx);   // the slot's reference; live decompressors hold their own
	g_ctx = ctx;
	return LZ4ADA_OK;
}

int lz4ada_init(int *min_buffer_size, int reservation, lz4ada_decompressor **out)   // lib/lz4ada.adb:48-63
{
	if (!min_buffer_size || !out || reservation < LZ4ADA_SZ_64_KIB || reservation > LZ4ADA_SZ_8_MIB)
		return LZ4ADA_ASSERTION_ERROR;
	const int block_max = block_size_of(reservation);
	*min_buffer_size = block_max + kHistorySize + 8;
	Meta m;
	m.reservation = reservation;
	*out = new lz4ada_decompressor(m, block_max + 4 + kBlockSizeBytes - 1, *min_buffer_size);
	return LZ4ADA_OK;
}

int lz4ada_init_with_header(const uint8_t *input, int input_len, int *num_consumed, int *min_buffer_size,
			    int reservation, lz4ada_decompressor **out, char *message, size_t message_cap)
{   // lib/lz4ada.adb:79-125
	if (!num_consumed || !min_buffer_size || !out) return LZ4ADA_ASSERTION_ERROR;
	*out = nullptr;
	*num_consumed = 0;
	Raised r;
	if (!input || input_len < 7) {   // Pre => Input'Length >= 7, lib/lz4ada.ads:243
		r = err_assertion("failed precondition from lz4ada.ads:243");
		copy_message(message, message_cap, r.text);
		return r.kind;
	}
	Meta mt;
	int in_last = 0;
	r = init_with_header_meta(input, input_len, reservation, mt, *num_consumed, in_last, *min_buffer_size);
	if (r) {
		copy_message(message, message_cap, r.text);
		return r.kind;
	}
	*out = new lz4ada_decompressor(mt, in_last, *min_buffer_size);
	return LZ4ADA_OK;
}

int lz4ada_init_for_block(int *min_buffer_size, int compressed_length, int reservation,
			  lz4ada_decompressor **out)   // lib/lz4ada.adb:127-147
{
	if (!min_buffer_size || !out || reservation < LZ4ADA_SZ_64_KIB || reservation > LZ4ADA_SZ_8_MIB)
		return LZ4ADA_ASSERTION_ERROR;
	const int block_max = block_size_of(reservation);
	// Input_Buffer is (0 .. Block_Max_Size - 1): a longer block cannot be cached (the reference dies with
	// Constraint_Error inside Cache_Data_And_Process_If_Full, lib/lz4ada.adb:646); a negative length is no length
	if (compressed_length < 0 || compressed_length > block_max) return LZ4ADA_ASSERTION_ERROR;
	*min_buffer_size = block_max + kHistorySize + 8;
	Meta m;
	m.format = Format::Block;
	m.is_compressed = true;
	m.stage = HeaderStage::Complete;
	m.reservation = reservation;
	lz4ada_decompressor *d = new lz4ada_decompressor(m, block_max - 1, *min_buffer_size);
	d->walker.input_length = compressed_length;
	*out = d;
	return LZ4ADA_OK;
}

int lz4ada_update(lz4ada_decompressor *ctx, const uint8_t *input, int input_len, int *num_consumed,
		  uint8_t *buffer, int buffer_len, int *output_first, int *output_last)
{
	if (!ctx || !num_consumed || !output_first || !output_last) return LZ4ADA_ASSERTION_ERROR;
	ctx->message.clear();
	Raised r = ctx->walker.update(input, input_len, *num_consumed, buffer, buffer_len, *output_first, *output_last);
	if (r) ctx->message = r.text;
	return r.kind;
}

int lz4ada_is_end_of_frame(const lz4ada_decompressor *ctx) { return ctx->walker.is_end_of_frame(); }

const char *lz4ada_exception_message(const lz4ada_decompressor *ctx) { return ctx ? ctx->message.c_str() : ""; }

void lz4ada_free(lz4ada_decompressor *ctx) { delete ctx; }

void lz4ada_to_hex_u8(uint8_t num, char *out)   // lib/lz4ada.adb:363-368
{
	static const char tbl[] = "0123456789abcdef";
	out[0] = tbl[num >> 4];
	out[1] = tbl[num & 15];
	out[2] = 0;
}

void lz4ada_to_hex_u32(uint32_t num, char *out)   // lib/lz4ada.adb:370-375
{
	for (int i = 0; i < 4; i++) lz4ada_to_hex_u8(uint8_t(num >> (24 - 8 * i)), out + 2 * i);
}

// XXHash32.Init ignores its Seed argument (lib/lz4ada.adb:925-930 calls Reset without it).
void lz4ada_xxhash32_init(lz4ada_xxhash32 *h, uint32_t seed)
{
	(void)seed;
	Xxh32Host::reset(h, 0);
}
void lz4ada_xxhash32_reset(lz4ada_xxhash32 *h, uint32_t seed) { Xxh32Host::reset(h, seed); }
void lz4ada_xxhash32_update(lz4ada_xxhash32 *h, const uint8_t *input, size_t len) { Xxh32Host::update(h, input, len); }
uint32_t lz4ada_xxhash32_final(const lz4ada_xxhash32 *h) { return Xxh32Host::final(h); }
uint32_t lz4ada_xxhash32_hash(const uint8_t *input, size_t len) { return Xxh32Host::hash(input, len); }

}  // extern "C"
