// common.hpp -- shared pieces of the host layer (the part the north star keeps in Ada; written
// in C++ here because this image has no Ada compiler -- ada/ holds the Ada rendering).
//
//   Raised            an LZ4Ada exception: kind + the exact GNAT Exception_Information line
//   Xxh32Host         package XXHash32 (lib/lz4ada.ads:311-344) for header checksums / public API
//   Meta, Walker      the reference's frame state machine (lib/lz4ada.adb:155-659)
//   BlockEngine       what happens to a complete block: decode on the device now (streaming) or
//                     record it in a block table (batch planner)
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "lz4b200.h"

namespace lz4ada {

constexpr int kHistorySize = 65536;      // lib/lz4ada.ads:350
constexpr int kBlockSizeBytes = 4;       // lib/lz4ada.ads:351
constexpr uint32_t kMagicModern = 0x184d2204u;   // lib/lz4ada.ads:348
constexpr uint32_t kMagicLegacy = 0x184c2102u;   // lib/lz4ada.ads:349
constexpr uint32_t kMagicSkipLo = 0x184d2a50u;   // lib/lz4ada.ads:353
constexpr uint32_t kMagicSkipHi = 0x184d2a5fu;

// An exception in flight.  kind = enum lz4ada_exception.
struct Raised {
	int kind = LZ4ADA_OK;
	std::string text;   // "raised LZ4ADA.<NAME> : <message>"
	explicit operator bool() const { return kind != LZ4ADA_OK; }
};

Raised raise(int kind, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
inline Raised ok() { return Raised(); }

// The reference's messages, one function per raise site.
Raised err_too_few_header_bytes(uint64_t remaining);                       // lib/lz4ada.adb:104
Raised err_bad_magic(uint32_t magic);                                      // :220
Raised err_too_little_memory(int effective, int requested);                // :246
Raised err_bad_version(unsigned version);                                  // :304
Raised err_reserved_bits();                                                // :310
Raised err_bad_block_max(unsigned code);                                   // :324
Raised err_header_checksum(unsigned computed, unsigned stored);            // :356
Raised err_single_frame_trailing();                                        // :439
Raised err_content_size_left(uint64_t remaining);                          // :471
Raised err_content_checksum(uint32_t computed, uint32_t declared);         // :505
Raised err_block_too_long(int buffer_len, uint32_t length, int metadata);  // :544
Raised err_single_frame_next_frame();                                      // :573
Raised err_block_checksum(uint32_t declared, uint32_t computed);           // :702
Raised err_ends_after_literals(int nibble);                                // :754
Raised err_offset_zero();                                                  // :770
Raised err_content_size_exceeded();                                        // :831
Raised err_backref_range(int value);                                       // :868
Raised err_library_bug();                                                  // :185
// Not in the reference (Appendix C: its behaviour there is undefined / Constraint_Error)
Raised err_literal_overrun(long long run, int left);
Raised err_length_ext_overrun(bool match);
Raised err_offset_truncated();
Raised err_output_exhausted(int buffer_len);
Raised err_block_length_unrepresentable(uint32_t length);
Raised err_block_exceeds_input_buffer(int buffer_len);
Raised err_assertion(const char *what);
Raised err_device(const char *what);

// Translate a device block status into the reference's exception (or ok()).
// `buffer_len` feeds the output-exhausted message.
Raised status_to_raised(const lz4b200_blk_status &st, int buffer_len);

int block_size_of(int reservation);   // Get_Block_Size, lib/lz4ada.adb:65-77

inline uint32_t load32(const uint8_t *p)
{
	return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}
inline uint64_t load64(const uint8_t *p) { return uint64_t(load32(p)) | (uint64_t(load32(p + 4)) << 32); }

// ---- package XXHash32 (host) -------------------------------------------------------------
struct Xxh32Host {
	static void reset(lz4ada_xxhash32 *h, uint32_t seed);
	static void update(lz4ada_xxhash32 *h, const uint8_t *p, size_t n);
	static uint32_t final(const lz4ada_xxhash32 *h);
	static uint32_t hash(const uint8_t *p, size_t n);
};

// ---- frame state machine -----------------------------------------------------------------
enum class Format { TBD, Legacy, Modern, Block, Skippable };                 // lib/lz4ada.ads:355
enum class HeaderStage { NeedMagic, NeedModern, NeedFlags, NeedSkippableLength, Complete };  // :356

struct Meta {   // Decompressor_Meta, lib/lz4ada.ads:359-370
	Format format = Format::TBD;
	HeaderStage stage = HeaderStage::NeedMagic;
	int reservation = LZ4ADA_FOR_ALL;
	int content_checksum_length = 0;
	int block_checksum_length = 0;
	int status_eof = LZ4ADA_EOF_NO;
	int input_buffer_filled = 0;
	bool is_compressed = false;
	bool has_content_size = false;
	bool block_independent = false;   // FLG bit 5 -- the reference ignores it; the planner does not
	int frame_block_max = 0;          // the frame's own maximum block size in bytes (BD code / legacy)
	uint64_t size_remaining = 4;
};

// Feeds header bytes (Process_Header_Bytes, lib/lz4ada.adb:155-191).
Raised header_feed(Meta &m, uint8_t *header_buffer, const uint8_t *input, int input_len, int &consumed);
Raised header_magic(Meta &m, uint32_t magic);   // Process_Header_Magic, :199-223

// Init_With_Header (lib/lz4ada.adb:79-125) up to the point where the Decompressor record is built: the
// first frame header parsed into `mt`, In_Last and Min_Buffer_Size as the reference computes them.
Raised init_with_header_meta(const uint8_t *input, int input_len, int reservation, Meta &mt, int &num_consumed,
			     int &in_last, int &min_buffer_size);

class Walker;

struct BlockEngine {
	virtual ~BlockEngine() {}
	// Reset_Outer_For_Next_Frame, lib/lz4ada.adb:451-461
	virtual Raised new_frame(Walker &w) = 0;
	// a frame header has just been completed (planner records it; streaming: nothing)
	virtual void frame_started(Walker &w) { (void)w; }
	// Decode_Full_Block_With_Trailer, :661-696.  blk = payload + optional 4-byte trailer.
	virtual Raised block(Walker &w, const uint8_t *blk, int blk_len, uint8_t *buffer, int buffer_len,
			     int &out_first, int &out_last) = 0;
	// end mark reached: compare the content checksum (:500-511), then content size left (:469-476)
	virtual Raised content_checksum(Walker &w, uint32_t declared) = 0;
	virtual Raised frame_ended(Walker &w) = 0;
	// skippable frame fully skipped / legacy frame interrupted by the next magic
	virtual void frame_closed(Walker &w) { (void)w; }
};

// The reference's Update step function (lib/lz4ada.adb:383-418 and everything it calls that is
// not block decoding).  One step per call; byte-granular resumability.
class Walker {
public:
	Walker(const Meta &m, int in_last, BlockEngine *engine);
	Raised update(const uint8_t *input, int input_len, int &consumed, uint8_t *buffer, int buffer_len,
		      int &out_first, int &out_last);
	int is_end_of_frame() const;   // lib/lz4ada.adb:906-915

	Meta m;
	bool at_end_mark = false;
	int input_length = -1;              // declared length of the current block, -1 = unknown
	int block_end_consumed = 0;         // value of `consumed` right after the block handed to the engine
	int input_buffer_len;               // Input_Buffer'Length = In_Last + 1 (what the limit checks see)
	std::vector<uint8_t> input_buffer;  // Input_Buffer(0 .. In_Last), grown on demand up to that length
	BlockEngine *engine;
	// While a block is handed to the engine straight out of the caller's Input (no copy): the rest of that
	// Input behind the block -- what the caller will present again next time.  Null otherwise.
	const uint8_t *lookahead = nullptr;
	int lookahead_len = 0;
	uint8_t *cache(size_t need);        // storage for the first `need` bytes of Input_Buffer

private:
	Raised skip(const uint8_t *input, int input_len, int &consumed);
	Raised reset_for_next_frame(const uint8_t *input, int input_len, int &consumed);
	Raised check_end_mark(const uint8_t *input, int input_len, int &consumed);
	Raised try_detect_input_length(const uint8_t *input, int input_len, int &consumed);
	Raised handle_newly_known_input_length(const uint8_t *input, int input_len, int &consumed, uint8_t *buffer,
					       int buffer_len, int &of, int &ol);
	Raised cache_data_and_process_if_full(const uint8_t *input, int input_len, int &consumed, uint8_t *buffer,
					      int buffer_len, int &of, int &ol);
	Raised header_bytes(const uint8_t *input, int input_len, int &consumed);
};

// process-wide default device context (lz4ada_set_device_context).  Returned with a reference taken
// for the caller (drop it with lz4b200_destroy).
lz4b200_ctx *default_context(Raised *why);

// One host-layer object attached to a device context and freed with it (shim.cu): made on first use.
void *ctx_attachment(lz4b200_ctx *ctx, void *(*make)(lz4b200_ctx *), void (*free_fn)(lz4b200_ctx *, void *));
// ... and a second one (the batch scratch pool has the first, the streaming engine's buffer pool this one)
void *ctx_attachment2(lz4b200_ctx *ctx, void *(*make)(lz4b200_ctx *), void (*free_fn)(lz4b200_ctx *, void *));

}  // namespace lz4ada
