// errors.cpp -- the reference's exception texts, formatted the way GNAT
// prints them: 'Image of a non-negative number carries a leading blank, hex is lower-case and
// zero padded (To_Hex, lib/lz4ada.adb:363-375), and the line starts with
// "raised LZ4ADA.<EXCEPTION> : " (test_suite/lz4test.adb:310-323 compares the whole line).
#include "common.hpp"

namespace lz4ada {

static const char *const kKindName[] = {"",
					"LZ4ADA.CHECKSUM_ERROR",
					"LZ4ADA.DATA_CORRUPTION",
					"LZ4ADA.NOT_SUPPORTED",
					"LZ4ADA.TOO_FEW_HEADER_BYTES",
					"LZ4ADA.TOO_LITTLE_MEMORY",
					"CONSTRAINT_ERROR",
					"ADA.ASSERTIONS.ASSERTION_ERROR",
					"LZ4ADA.DEVICE_ERROR"};

// Flexible_Memory_Reservation'Image, lib/lz4ada.ads:79-80
static const char *const kReservationImage[] = {"SZ_64_KIB", "SZ_256_KIB", "SZ_1_MIB", "SZ_4_MIB",
						"SZ_8_MIB",  "USE_FIRST",  "SINGLE_FRAME"};

Raised raise(int kind, const char *fmt, ...)
{
	char body[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(body, sizeof body, fmt, ap);
	va_end(ap);
	Raised r;
	r.kind = kind;
	r.text = std::string("raised ") + kKindName[kind] + " : " + body;
	return r;
}

// Integer'Image / U64'Image: blank in front of non-negative values
static std::string image(long long v)
{
	char b[32];
	snprintf(b, sizeof b, v < 0 ? "%lld" : " %lld", v);
	return b;
}
static std::string image_u(unsigned long long v)
{
	char b[32];
	snprintf(b, sizeof b, " %llu", v);
	return b;
}

Raised err_too_few_header_bytes(uint64_t remaining)
{
	return raise(LZ4ADA_TOO_FEW_HEADER_BYTES,
		     "Expected at least %s more bytes but header input has already ended.", image_u(remaining).c_str());
}
Raised err_bad_magic(uint32_t magic)
{
	return raise(LZ4ADA_NOT_SUPPORTED, "Invalid or unsupported magic: 0x%08x", magic);
}
Raised err_too_little_memory(int effective, int requested)
{
	return raise(LZ4ADA_TOO_LITTLE_MEMORY,
		     "LZ4 header requres reservation %s, but API call requested that only %s be used. "
		     "This frame cannot be processed under the given constraints.",
		     kReservationImage[effective], kReservationImage[requested]);
}
Raised err_bad_version(unsigned version)
{
	return raise(LZ4ADA_NOT_SUPPORTED, "Only LZ4 frame format version 01 supported. Detected 0x%02x instead.",
		     version);
}
Raised err_reserved_bits()
{
	return raise(LZ4ADA_NOT_SUPPORTED,
		     "Found reserved bits /= 0. Data might be too new to be processed by this implementation!");
}
Raised err_bad_block_max(unsigned code)
{
	return raise(LZ4ADA_NOT_SUPPORTED, "Unknown maximum block size flag: 0x%02x", code);
}
Raised err_header_checksum(unsigned computed, unsigned stored)
{
	return raise(LZ4ADA_CHECKSUM_ERROR,
		     "Computed Header Checksum 0x%02x does not match expected Header Checksum 0x%02x", computed,
		     stored);
}
Raised err_single_frame_trailing()
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Requested Single_Frame operation but data was provided after End of Frame was detected");
}
Raised err_content_size_left(uint64_t remaining)
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Frame has ended, but according to content size, there should be %s bytes left to output.",
		     image_u(remaining).c_str());
}
Raised err_content_checksum(uint32_t computed, uint32_t declared)
{
	return raise(LZ4ADA_CHECKSUM_ERROR,
		     "Computed content checksum 0x%08x does not match declared content checksum 0x%08x.", computed,
		     declared);
}
Raised err_block_too_long(int buffer_len, uint32_t length, int metadata)
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Declared maximum data length exceeded. Buffer has %s bytes, current block requires %s bytes + "
		     "%s bytes for metadata.",
		     image(buffer_len).c_str(), image_u(length).c_str(), image(metadata).c_str());
}
Raised err_single_frame_next_frame()
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Requested Single_Frame operation but data provided what looks like the beginning of another "
		     "frame.");
}
Raised err_block_checksum(uint32_t declared, uint32_t computed)
{
	return raise(LZ4ADA_CHECKSUM_ERROR, "Declared checksum is 0x%08x, but computed one is 0x%08x.", declared,
		     computed);
}
Raised err_ends_after_literals(int nibble)
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Match_Length=%s suggests compressed data but this sequence already ends after the literals. "
		     "This might also happen with an untypical encoder?",
		     image(nibble).c_str());
}
Raised err_offset_zero() { return raise(LZ4ADA_DATA_CORRUPTION, "Corrupted Block: Offset = 0 detected."); }
Raised err_content_size_exceeded()
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Produced content size exceeds declared content size. The supplied data is inconsistent.");
}
Raised err_backref_range(int value)
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Backreference location out of range. Read from offset %s not possible (earliest available "
		     "index is 0).",
		     image(value).c_str());
}
Raised err_library_bug()
{
	return raise(LZ4ADA_CONSTRAINT_ERROR,
		     "Header_Complete case must not be reached while processing header bytes. Library bug detected.");
}
Raised err_literal_overrun(long long run, int left)
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Corrupted Block: Literal run of %lld bytes exceeds the %d bytes left in the block.", run, left);
}
Raised err_length_ext_overrun(bool match)
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Corrupted Block: %s length extension runs past the end of the block.",
		     match ? "Match" : "Literal");
}
Raised err_offset_truncated()
{
	return raise(LZ4ADA_DATA_CORRUPTION, "Corrupted Block: Block ends inside a match offset.");
}
Raised err_output_exhausted(int buffer_len)
{
	return raise(LZ4ADA_DATA_CORRUPTION,
		     "Output buffer exhausted. Decompressed data does not fit into the %d bytes provided.", buffer_len);
}
Raised err_block_length_unrepresentable(uint32_t length)
{
	return raise(LZ4ADA_DATA_CORRUPTION, "Declared block length %u is not representable.", length);
}
Raised err_block_exceeds_input_buffer(int buffer_len)
{
	return raise(LZ4ADA_DATA_CORRUPTION, "Declared block length exceeds the input buffer of %s bytes.",
		     image(buffer_len).c_str());
}
Raised err_assertion(const char *what) { return raise(LZ4ADA_ASSERTION_ERROR, "%s", what); }
Raised err_device(const char *what) { return raise(LZ4ADA_DEVICE_ERROR, "%s", what); }

Raised status_to_raised(const lz4b200_blk_status &st, int buffer_len)
{
	switch (st.code) {
	case LZ4B200_ST_OK: return ok();
	case LZ4B200_ST_BLOCK_CHECKSUM: return err_block_checksum(st.xxh32_declared, st.xxh32_computed);
	case LZ4B200_ST_ENDS_AFTER_LITERALS: return err_ends_after_literals(st.aux);
	case LZ4B200_ST_OFFSET_ZERO: return err_offset_zero();
	case LZ4B200_ST_BACKREF_RANGE: return err_backref_range(st.aux);
	case LZ4B200_ST_LITERAL_OVERRUN: return err_literal_overrun(st.out_len, st.aux);
	case LZ4B200_ST_LIT_EXT_OVERRUN: return err_length_ext_overrun(false);
	case LZ4B200_ST_MATCH_EXT_OVERRUN: return err_length_ext_overrun(true);
	case LZ4B200_ST_OFFSET_TRUNCATED: return err_offset_truncated();
	case LZ4B200_ST_OUTPUT_OVERFLOW: return err_output_exhausted(buffer_len);
	default: return err_device("unexpected block status from the device");
	}
}

int block_size_of(int reservation)
{
	static const int lut[] = {64 * 1024, 256 * 1024, 1024 * 1024, 4 * 1024 * 1024, 8 * 1024 * 1024};
	return lut[reservation];
}

}  // namespace lz4ada
