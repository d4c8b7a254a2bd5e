// walker.cpp -- frame-header parser and the Update step function of the host layer.
//
// Behavioural twin of lib/lz4ada.adb:155-361 (header) and :383-659 (stream state machine): the
// same stages, the same "one step per call" granularity, the same exceptions in the same order
// (SURVEY.md Appendix A / B).  Block payloads are never touched here -- a complete block is
// handed to the BlockEngine, which sends it to the GPU (streaming) or records it (batch planner).
#include "common.hpp"

namespace lz4ada {

// Check_Reservation, lib/lz4ada.adb:241-260: a fixed request caps the frame's need and then
// *replaces* it (the input buffer was sized for the request).
static Raised check_reservation(int requested, int &effective)
{
	if (requested <= LZ4ADA_SZ_8_MIB) {
		if (effective > requested) return err_too_little_memory(effective, requested);
		effective = requested;
	}
	return ok();
}

static Raised legacy_header_done(Meta &m)   // Process_Legacy_End_Of_Header, :225-239
{
	int effective = LZ4ADA_FOR_LEGACY;
	m.input_buffer_filled = 0;
	m.format = Format::Legacy;
	m.stage = HeaderStage::Complete;
	m.size_remaining = 0;
	m.status_eof = LZ4ADA_EOF_MAYBE;
	m.block_checksum_length = 0;
	m.content_checksum_length = 0;
	m.has_content_size = false;
	m.is_compressed = true;
	m.block_independent = true;   // legacy blocks are compressed one by one
	m.frame_block_max = 8 * 1024 * 1024;
	if (Raised r = check_reservation(m.reservation, effective)) return r;
	m.reservation = effective;
	return ok();
}

Raised header_magic(Meta &m, uint32_t magic)   // Process_Header_Magic, :199-223
{
	if (magic == kMagicModern) {
		m.format = Format::Modern;
		m.stage = HeaderStage::NeedFlags;
		m.size_remaining = 2;
		return ok();
	}
	if (magic == kMagicLegacy) return legacy_header_done(m);
	if (magic >= kMagicSkipLo && magic <= kMagicSkipHi) {
		m.format = Format::Skippable;
		m.stage = HeaderStage::NeedSkippableLength;
		m.size_remaining = 4;
		m.block_checksum_length = 0;
		m.content_checksum_length = 0;
		return ok();
	}
	return err_bad_magic(magic);
}

static Raised header_flags(Meta &m, const uint8_t *hb)   // Process_Header_Flags, :262-298
{
	const uint8_t flg = hb[4], bd = hb[5];
	const unsigned version = (flg & 0xc0u) >> 6;
	if (version != 1) return err_bad_version(version);                       // :303
	if ((flg & 0x02u) || (bd & 0x8fu)) return err_reserved_bits();           // :309
	m.status_eof = LZ4ADA_EOF_NO;
	const unsigned code = (bd & 0x70u) >> 4;
	int required;
	switch (code) {                                                           // :316-328
	case 4: required = LZ4ADA_SZ_64_KIB; break;
	case 5: required = LZ4ADA_SZ_256_KIB; break;
	case 6: required = LZ4ADA_SZ_1_MIB; break;
	case 7: required = LZ4ADA_SZ_4_MIB; break;
	default: return err_bad_block_max(code);
	}
	m.block_checksum_length = (flg & 0x10u) ? 4 : 0;
	m.content_checksum_length = (flg & 0x04u) ? 4 : 0;
	m.has_content_size = (flg & 0x08u) != 0;
	m.block_independent = (flg & 0x20u) != 0;
	m.frame_block_max = block_size_of(required);
	m.stage = HeaderStage::NeedModern;
	// descriptor tail still to come: HC byte, optional content size, optional dictionary id
	m.size_remaining = 1 + (m.has_content_size ? 8 : 0) + ((flg & 0x01u) ? 4 : 0);
	if (Raised r = check_reservation(m.reservation, required)) return r;
	if (m.reservation != LZ4ADA_SINGLE_FRAME) m.reservation = required;
	return ok();
}

static Raised modern_header_done(Meta &m, const uint8_t *hb)   // Process_Modern_End_Of_Header, :330-361
{
	const int filled = m.input_buffer_filled;
	const uint8_t stored = hb[filled - 1];
	if (m.has_content_size) m.size_remaining = load64(hb + 6);
	const uint8_t computed = uint8_t((Xxh32Host::hash(hb + 4, size_t(filled - 5)) >> 8) & 0xffu);
	if (stored != computed) return err_header_checksum(computed, stored);
	m.stage = HeaderStage::Complete;
	m.input_buffer_filled = 0;
	return ok();
}

Raised header_feed(Meta &m, uint8_t *hb, const uint8_t *input, int input_len, int &consumed)
{
	int n = input_len;
	if (uint64_t(n) > m.size_remaining) n = int(m.size_remaining);
	if (n <= 0) return err_assertion("Copy_Length > 0");   // Ada.Assertions.Assert, :161
	memcpy(hb + m.input_buffer_filled, input, size_t(n));
	m.input_buffer_filled += n;
	m.size_remaining -= uint64_t(n);
	consumed = n;
	if (m.size_remaining != 0) return ok();
	switch (m.stage) {
	case HeaderStage::NeedMagic: return header_magic(m, load32(hb));
	case HeaderStage::NeedFlags: return header_flags(m, hb);
	case HeaderStage::NeedModern: return modern_header_done(m, hb);
	case HeaderStage::NeedSkippableLength:
		m.reservation = LZ4ADA_SZ_64_KIB;   // :177 (sic -- later frames are checked against it)
		m.stage = HeaderStage::Complete;
		m.size_remaining = load32(hb + 4);
		m.status_eof = m.size_remaining == 0 ? LZ4ADA_EOF_YES : LZ4ADA_EOF_NO;
		m.input_buffer_filled = 0;
		return ok();
	default: return err_library_bug();
	}
}

// Init_With_Header, lib/lz4ada.adb:79-125: feed header bytes until the first header is complete.
Raised init_with_header_meta(const uint8_t *input, int input_len, int reservation, Meta &mt, int &num_consumed,
			     int &in_last, int &min_buffer_size)
{
	num_consumed = 0;
	uint8_t header_buffer[20];
	mt = Meta();
	mt.reservation = reservation == LZ4ADA_SINGLE_FRAME ? LZ4ADA_USE_FIRST : reservation;
	int pos = 0;
	while (mt.stage != HeaderStage::Complete) {
		if (pos >= input_len) return err_too_few_header_bytes(mt.size_remaining);   // :104
		int inner = 0;
		if (Raised r = header_feed(mt, header_buffer, input + pos, input_len - pos, inner)) return r;
		pos += inner;
		num_consumed += inner;
	}
	const int block_max = block_size_of(mt.reservation);
	in_last = block_max + mt.block_checksum_length + kBlockSizeBytes - 1;   // :117-118
	min_buffer_size = block_max + kHistorySize + 8;                          // :119
	if (reservation == LZ4ADA_SINGLE_FRAME) mt.reservation = LZ4ADA_SINGLE_FRAME;
	return ok();
}

// ------------------------------------------------------------------------------------------

Walker::Walker(const Meta &meta, int in_last, BlockEngine *eng)
	: m(meta), input_buffer_len(in_last + 1), input_buffer(64, 0), engine(eng)
{
}

// The reference embeds Input_Buffer(0 .. In_Last) in the context (up to 8 MiB + 8, zeroed);
// only header bytes and blocks that arrive in pieces ever live there, so grow on demand.
uint8_t *Walker::cache(size_t need)
{
	if (need > input_buffer.size()) {
		size_t n = input_buffer.size();
		while (n < need) n *= 2;
		if (n > size_t(input_buffer_len) + 32) n = size_t(input_buffer_len) + 32;
		if (n < need) n = need;
		input_buffer.resize(n, 0);
	}
	return input_buffer.data();
}

int Walker::is_end_of_frame() const
{
	switch (m.format) {
	case Format::Legacy: return at_end_mark ? LZ4ADA_EOF_MAYBE : m.status_eof;
	case Format::Block: return input_length == -1 ? LZ4ADA_EOF_YES : LZ4ADA_EOF_NO;
	default: return m.status_eof;
	}
}

Raised Walker::header_bytes(const uint8_t *input, int input_len, int &consumed)
{
	const bool was_incomplete = m.stage != HeaderStage::Complete;
	Raised r = header_feed(m, cache(64), input, input_len, consumed);
	if (!r && was_incomplete && m.stage == HeaderStage::Complete) engine->frame_started(*this);
	return r;
}

Raised Walker::update(const uint8_t *input, int input_len, int &consumed, uint8_t *buffer, int buffer_len,
		      int &of, int &ol)
{
	consumed = 0;
	of = 1;
	ol = 0;
	if (m.stage != HeaderStage::Complete) return header_bytes(input, input_len, consumed);
	if (m.format == Format::Skippable) return skip(input, input_len, consumed);
	if (m.format == Format::TBD) return err_assertion("Is_Format /= TBD");
	if (at_end_mark) return check_end_mark(input, input_len, consumed);
	if (input_length != -1) return cache_data_and_process_if_full(input, input_len, consumed, buffer, buffer_len, of, ol);
	if (Raised r = try_detect_input_length(input, input_len, consumed)) return r;
	if (at_end_mark) return check_end_mark(input, input_len, consumed);
	if (input_length != -1) return handle_newly_known_input_length(input, input_len, consumed, buffer, buffer_len, of, ol);
	return ok();
}

Raised Walker::skip(const uint8_t *input, int input_len, int &consumed)   // :420-433
{
	const uint64_t remain = m.size_remaining;
	const uint64_t take = uint64_t(input_len) < remain ? uint64_t(input_len) : remain;
	if (m.status_eof == LZ4ADA_EOF_YES && take == 0) return reset_for_next_frame(input, input_len, consumed);
	consumed = int(take);
	m.size_remaining = remain - take;
	m.status_eof = m.size_remaining == 0 ? LZ4ADA_EOF_YES : LZ4ADA_EOF_NO;
	if (m.size_remaining == 0) engine->frame_closed(*this);
	return ok();
}

Raised Walker::reset_for_next_frame(const uint8_t *input, int input_len, int &consumed)   // :435-449
{
	if (m.reservation == LZ4ADA_SINGLE_FRAME) return err_single_frame_trailing();
	m.status_eof = LZ4ADA_EOF_NO;
	m.stage = HeaderStage::NeedMagic;
	m.size_remaining = 4;
	at_end_mark = false;
	input_length = -1;
	if (Raised r = engine->new_frame(*this)) return r;
	return header_bytes(input, input_len, consumed);
}

Raised Walker::check_end_mark(const uint8_t *input, int input_len, int &consumed)   // :463-523
{
	const int provided = input_len - consumed;
	const int required = m.content_checksum_length - m.input_buffer_filled;
	if (m.content_checksum_length == 0 || m.status_eof == LZ4ADA_EOF_YES || required <= 0) {
		if (m.status_eof == LZ4ADA_EOF_YES) {
			if (consumed != 0) return err_assertion("Num_Consumed = 0");
			return reset_for_next_frame(input, input_len, consumed);
		}
	} else if (provided >= required) {
		uint8_t word[4];
		memcpy(word, cache(8), size_t(m.input_buffer_filled));
		memcpy(word + m.input_buffer_filled, input + consumed, size_t(required));
		consumed += required;
		if (Raised r = engine->content_checksum(*this, load32(word))) return r;
	} else {
		memcpy(cache(8) + m.input_buffer_filled, input + consumed, size_t(provided));
		m.input_buffer_filled += provided;
		consumed += provided;
		return ok();
	}
	// Set_Frame_Has_Ended, :465-477
	m.status_eof = LZ4ADA_EOF_YES;
	m.input_buffer_filled = 0;
	return engine->frame_ended(*this);
}

static bool is_any_magic(uint32_t v)   // :587-593
{
	return v == kMagicModern || v == kMagicLegacy || (v >= kMagicSkipLo && v <= kMagicSkipHi);
}

Raised Walker::try_detect_input_length(const uint8_t *input, int input_len, int &consumed)   // :525-585
{
	int take = kBlockSizeBytes - m.input_buffer_filled;
	if (take > input_len) take = input_len;
	consumed = take;
	memcpy(cache(8) + m.input_buffer_filled, input, size_t(take));
	m.input_buffer_filled += take;
	if (m.input_buffer_filled != kBlockSizeBytes) return ok();

	uint32_t word = load32(cache(8));
	if (m.format == Format::Modern && word == 0) {
		at_end_mark = true;
		m.input_buffer_filled = 0;
		return ok();
	}
	if (m.format == Format::Legacy && is_any_magic(word)) {
		if (m.reservation == LZ4ADA_SINGLE_FRAME) return err_single_frame_next_frame();
		engine->frame_closed(*this);
		at_end_mark = false;
		input_length = -1;
		if (Raised r = engine->new_frame(*this)) return r;
		const bool modern_or_skip = word != kMagicLegacy;
		Raised r = header_magic(m, word);
		if (!r && !modern_or_skip) engine->frame_started(*this);   // legacy header is complete already
		return r;
	}
	if (m.format == Format::Modern) {
		m.is_compressed = (word & 0x80000000u) == 0;
		word &= 0x7ffffffu;   // 27 bits, :538 (sic)
	}
	if (word > 0x7fffffffu) return err_block_length_unrepresentable(word);
	const int additional = kBlockSizeBytes + m.block_checksum_length;
	if (int64_t(word) + additional > int64_t(input_buffer_len)) {
		input_length = -1;
		return err_block_too_long(input_buffer_len, word, additional);
	}
	input_length = int(word);
	return ok();
}

Raised Walker::handle_newly_known_input_length(const uint8_t *input, int input_len, int &consumed,
					       uint8_t *buffer, int buffer_len, int &of, int &ol)   // :595-628
{
	const int total = input_length + m.block_checksum_length;
	if (input_len - consumed >= total) {
		const uint8_t *blk = input + consumed;   // whole block present: no copy
		consumed += total;
		block_end_consumed = consumed;
		m.input_buffer_filled = 0;
		input_length = -1;
		lookahead = input + consumed;
		lookahead_len = input_len - consumed;
		const Raised r = engine->block(*this, blk, total, buffer, buffer_len, of, ol);
		lookahead = nullptr;
		lookahead_len = 0;
		return r;
	}
	return cache_data_and_process_if_full(input, input_len, consumed, buffer, buffer_len, of, ol);
}

Raised Walker::cache_data_and_process_if_full(const uint8_t *input, int input_len, int &consumed,
					      uint8_t *buffer, int buffer_len, int &of, int &ol)   // :630-659
{
	const int avail = input_len - consumed;
	const int head = m.format == Format::Block ? 0 : kBlockSizeBytes;
	const int fill = m.input_buffer_filled;
	const int want = input_length + m.block_checksum_length - fill + head;
	const uint8_t *src = input + consumed;
	const int take = want > avail ? avail : want;
	if (int64_t(fill) + take > int64_t(input_buffer_len))
		return err_block_exceeds_input_buffer(input_buffer_len);
	memcpy(cache(size_t(fill) + size_t(take)) + fill, src, size_t(take));
	consumed += take;
	if (want > avail) {
		m.input_buffer_filled += avail;
		return ok();
	}
	m.input_buffer_filled = 0;
	input_length = -1;
	block_end_consumed = consumed;
	// the reference drops the first four cached bytes for the raw-block API too (:654);
	// here `head` is 0 for Format::Block (Appendix C)
	return engine->block(*this, input_buffer.data() + head, fill - head + want, buffer, buffer_len, of, ol);
}

}  // namespace lz4ada
