/*
 * lz4ada_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See lz4ada_oracle.h for the rules about who may use this file.
 *
 * Restates /root/reference/lib/lz4ada.adb.  Every function names the
 * reference lines it follows.  The reference keeps decoded data in a ring
 * inside the caller's buffer (Output_Pos / Output_Pos_History); the oracle
 * keeps the same cursors so that Output_First / Output_Last agree too.
 */
#include "lz4ada_oracle.h"

#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define HISTORY_SIZE     65536   /* lib/lz4ada.ads:350 */
#define BLOCK_SIZE_BYTES 4       /* lib/lz4ada.ads:351 */
#define MAGIC_MODERN     0x184d2204u  /* lib/lz4ada.ads:348 */
#define MAGIC_LEGACY     0x184c2102u  /* lib/lz4ada.ads:349 */
#define MAGIC_SKIP_LO    0x184d2a50u  /* lib/lz4ada.ads:353 */
#define MAGIC_SKIP_HI    0x184d2a5fu

/* lib/lz4ada.ads:355-357 */
enum format { F_TBD, F_LEGACY, F_MODERN, F_BLOCK, F_SKIPPABLE };
enum hstate { NEED_MAGIC, NEED_MODERN, NEED_FLAGS, NEED_SKIPPABLE_LENGTH,
	      HEADER_COMPLETE };

/* lib/lz4ada.ads:359-370 */
struct meta {
	int      is_format;
	int      header_parsing;
	int      memory_reservation;
	int      content_checksum_length;
	int      block_checksum_length;
	int      status_eof;
	int      input_buffer_filled;
	int      is_compressed;
	int      has_content_size;
	uint64_t size_remaining;
};

/* lib/lz4ada.ads:440-449 */
struct lzo_ctx {
	struct meta m;
	int      is_at_end_mark;
	uint8_t *input_buffer;
	int      input_buffer_len;      /* In_Last + 1 */
	int      output_pos;
	int      output_pos_history;
	int      input_length;
	lzo_xxh32 hash_all_data;
	char     msg[640];
};

static const char *const EXC_NAME[] = {
	"", "CHECKSUM_ERROR", "DATA_CORRUPTION", "NOT_SUPPORTED",
	"TOO_FEW_HEADER_BYTES", "TOO_LITTLE_MEMORY", "CONSTRAINT_ERROR",
	"ASSERTION_ERROR"
};

static const char *const RES_IMAGE[] = {  /* Flexible_Memory_Reservation'Image */
	"SZ_64_KIB", "SZ_256_KIB", "SZ_1_MIB", "SZ_4_MIB", "SZ_8_MIB",
	"USE_FIRST", "SINGLE_FRAME"
};

static int raise_to(char *msg, size_t cap, int exc, const char *fmt, ...)
{
	va_list ap;
	int n;
	if (msg == NULL || cap == 0)
		return exc;
	if (exc == LZO_CONSTRAINT_ERROR)
		n = snprintf(msg, cap, "raised CONSTRAINT_ERROR : ");
	else if (exc == LZO_ASSERTION_ERROR)
		n = snprintf(msg, cap, "raised ADA.ASSERTIONS.ASSERTION_ERROR : ");
	else
		n = snprintf(msg, cap, "raised LZ4ADA.%s : ", EXC_NAME[exc]);
	va_start(ap, fmt);
	vsnprintf(msg + n, cap - (size_t)n, fmt, ap);
	va_end(ap);
	return exc;
}

#define RAISE(m, exc, ...) return raise_to((m), 640, (exc), __VA_ARGS__)
#define TRY(expr) do { int rc_ = (expr); if (rc_ != LZO_OK) return rc_; } while (0)

static uint32_t load_32(const uint8_t *p)   /* lib/lz4ada.ads:451-456 */
{
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
	       ((uint32_t)p[3] << 24);
}

static uint64_t load_64(const uint8_t *p)   /* lib/lz4ada.adb:345-349 */
{
	return (uint64_t)load_32(p) | ((uint64_t)load_32(p + 4) << 32);
}

/* Get_Block_Size, lib/lz4ada.adb:65-77 */
static int get_block_size(int r)
{
	static const int lut[] = { 64 * 1024, 256 * 1024, 1024 * 1024,
				   4 * 1024 * 1024, 8 * 1024 * 1024 };
	return lut[r];
}

/* ------------------------------------------------------------ XXHash32 -- */

#define PRIME_1 2654435761u   /* lib/lz4ada.ads:324-328 */
#define PRIME_2 2246822519u
#define PRIME_3 3266489917u
#define PRIME_4  668265263u
#define PRIME_5  374761393u

static uint32_t rotl(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

void lzo_xxh32_reset(lzo_xxh32 *h, uint32_t seed)   /* :932-940 */
{
	h->state[0] = seed + PRIME_1 + PRIME_2;
	h->state[1] = seed + PRIME_2;
	h->state[2] = seed;
	h->state[3] = seed - PRIME_1;
	h->buffer_size = 0;
	h->total_length = 0;
}

static void xxh32_process(lzo_xxh32 *h, const uint8_t *d)   /* :979-991 */
{
	int i;
	for (i = 0; i < 4; i++)
		h->state[i] = rotl(h->state[i] + load_32(d + 4 * i) * PRIME_2,
				   13) * PRIME_1;
}

void lzo_xxh32_update(lzo_xxh32 *h, const uint8_t *p, size_t n)  /* :942-977 */
{
	size_t start = 0;
	int fast = (h->buffer_size == 0);
	while (start < n) {
		if (n - start >= 16 && fast) {
			h->total_length += 16;
			xxh32_process(h, p + start);
			start += 16;
		} else {                          /* Update1, :965-977 */
			h->buffer[h->buffer_size++] = p[start++];
			h->total_length += 1;
			fast = 0;
			if (h->buffer_size == 16) {
				h->buffer_size = 0;
				xxh32_process(h, h->buffer);
				fast = 1;
			}
		}
	}
}

uint32_t lzo_xxh32_final(const lzo_xxh32 *h)   /* :993-1017 */
{
	uint32_t ret = (uint32_t)(h->total_length & 0xffffffffu);
	int data = 0;
	if (h->total_length >= 16)
		ret += rotl(h->state[0], 1) + rotl(h->state[1], 7) +
		       rotl(h->state[2], 12) + rotl(h->state[3], 18);
	else
		ret += h->state[2] + PRIME_5;
	while (data + 3 < h->buffer_size) {
		ret = rotl(ret + load_32(h->buffer + data) * PRIME_3, 17) *
		      PRIME_4;
		data += 4;
	}
	while (data < h->buffer_size) {
		ret = rotl(ret + (uint32_t)h->buffer[data] * PRIME_5, 11) *
		      PRIME_1;
		data += 1;
	}
	ret = (ret ^ (ret >> 15)) * PRIME_2;
	ret = (ret ^ (ret >> 13)) * PRIME_3;
	return ret ^ (ret >> 16);
}

uint32_t lzo_xxh32_hash(const uint8_t *p, size_t n)   /* :1019-1024 */
{
	lzo_xxh32 h;
	lzo_xxh32_reset(&h, 0);
	lzo_xxh32_update(&h, p, n);
	return lzo_xxh32_final(&h);
}

/* -------------------------------------------------------------- Header -- */

/* Check_Reservation, lib/lz4ada.adb:241-260 */
static int check_reservation(char *msg, int requested, int *effective)
{
	if (requested <= LZO_SZ_8_MIB) {
		if (*effective > requested)
			RAISE(msg, LZO_TOO_LITTLE_MEMORY,
			      "LZ4 header requres reservation %s, but API call "
			      "requested that only %s be used. This frame "
			      "cannot be processed under the given constraints.",
			      RES_IMAGE[*effective], RES_IMAGE[requested]);
		*effective = requested;
	}
	return LZO_OK;
}

/* Process_Legacy_End_Of_Header, lib/lz4ada.adb:225-239 */
static int process_legacy_end_of_header(struct meta *m, char *msg)
{
	int effective = LZO_FOR_LEGACY;
	m->input_buffer_filled = 0;
	m->is_format = F_LEGACY;
	m->header_parsing = HEADER_COMPLETE;
	m->size_remaining = 0;
	m->status_eof = LZO_EOF_MAYBE;
	m->block_checksum_length = 0;
	m->content_checksum_length = 0;
	m->has_content_size = 0;
	m->is_compressed = 1;
	TRY(check_reservation(msg, m->memory_reservation, &effective));
	m->memory_reservation = effective;
	return LZO_OK;
}

/* Process_Header_Magic, lib/lz4ada.adb:199-223 */
static int process_header_magic(struct meta *m, char *msg, uint32_t magic)
{
	if (magic == MAGIC_MODERN) {
		m->is_format = F_MODERN;
		m->header_parsing = NEED_FLAGS;
		m->size_remaining = 2;
	} else if (magic == MAGIC_LEGACY) {
		TRY(process_legacy_end_of_header(m, msg));
	} else if (magic >= MAGIC_SKIP_LO && magic <= MAGIC_SKIP_HI) {
		m->is_format = F_SKIPPABLE;
		m->header_parsing = NEED_SKIPPABLE_LENGTH;
		m->size_remaining = 4;
		m->block_checksum_length = 0;
		m->content_checksum_length = 0;
	} else {
		RAISE(msg, LZO_NOT_SUPPORTED,
		      "Invalid or unsupported magic: 0x%08x", magic);
	}
	return LZO_OK;
}

/* Process_Header_Flags, lib/lz4ada.adb:262-298 with Check_Flag_Validity
 * (:300-314) and Get_Block_Size_Reservation (:316-328) folded in. */
static int process_header_flags(struct meta *m, char *msg, const uint8_t *ib)
{
	uint8_t flg = ib[4], bd = ib[5];
	int flg_version = (flg & 0xc0) >> 6;
	int reserved = ((flg & 2) != 0) || ((bd & 0x8f) != 0);
	int bd_block_max = (bd & 0x70) >> 4;
	int required;

	if (flg_version != 1)
		RAISE(msg, LZO_NOT_SUPPORTED,
		      "Only LZ4 frame format version 01 supported. "
		      "Detected 0x%02x instead.", flg_version);
	if (reserved)
		RAISE(msg, LZO_NOT_SUPPORTED,
		      "Found reserved bits /= 0. Data might be too new to be "
		      "processed by this implementation!");
	m->status_eof = LZO_EOF_NO;
	switch (bd_block_max) {
	case 4: required = LZO_SZ_64_KIB;  break;
	case 5: required = LZO_SZ_256_KIB; break;
	case 6: required = LZO_SZ_1_MIB;   break;
	case 7: required = LZO_SZ_4_MIB;   break;
	default:
		RAISE(msg, LZO_NOT_SUPPORTED,
		      "Unknown maximum block size flag: 0x%02x", bd_block_max);
	}
	m->block_checksum_length = (flg & 16) ? 4 : 0;
	m->content_checksum_length = (flg & 4) ? 4 : 0;
	m->has_content_size = (flg & 8) != 0;
	m->header_parsing = NEED_MODERN;
	m->size_remaining = (uint64_t)(1 + (m->has_content_size ? 8 : 0) +
				       ((flg & 1) ? 4 : 0));
	TRY(check_reservation(msg, m->memory_reservation, &required));
	if (m->memory_reservation != LZO_SINGLE_FRAME)
		m->memory_reservation = required;
	return LZO_OK;
}

/* Process_Modern_End_Of_Header (:330-343) + Check_Header_Checksum (:351-361) */
static int process_modern_end_of_header(struct meta *m, char *msg,
					const uint8_t *ib)
{
	uint8_t hc = ib[m->input_buffer_filled - 1];
	uint8_t computed;
	if (m->has_content_size)
		m->size_remaining = load_64(ib + 6);
	computed = (uint8_t)((lzo_xxh32_hash(ib + 4,
			(size_t)(m->input_buffer_filled - 1 - 4)) >> 8) & 0xff);
	if (hc != computed)
		RAISE(msg, LZO_CHECKSUM_ERROR,
		      "Computed Header Checksum 0x%02x does not match "
		      "expected Header Checksum 0x%02x", computed, hc);
	m->header_parsing = HEADER_COMPLETE;
	m->input_buffer_filled = 0;
	return LZO_OK;
}

/* Process_Header_Bytes, lib/lz4ada.adb:155-191 */
static int process_header_bytes(struct meta *m, char *msg, uint8_t *ib,
				const uint8_t *input, int input_len,
				int *num_consumed)
{
	int copy_length = input_len;
	if ((uint64_t)copy_length > m->size_remaining)
		copy_length = (int)m->size_remaining;
	if (copy_length <= 0)   /* Ada.Assertions.Assert(Copy_Length > 0), :161 */
		RAISE(msg, LZO_ASSERTION_ERROR, "Copy_Length > 0");
	memcpy(ib + m->input_buffer_filled, input, (size_t)copy_length);
	m->input_buffer_filled += copy_length;
	m->size_remaining -= (uint64_t)copy_length;
	*num_consumed = copy_length;
	if (m->size_remaining == 0) {
		switch (m->header_parsing) {
		case NEED_MAGIC:
			TRY(process_header_magic(m, msg, load_32(ib)));
			break;
		case NEED_FLAGS:
			TRY(process_header_flags(m, msg, ib));
			break;
		case NEED_MODERN:
			TRY(process_modern_end_of_header(m, msg, ib));
			break;
		case NEED_SKIPPABLE_LENGTH:
			m->memory_reservation = LZO_SZ_64_KIB;   /* :177 */
			m->header_parsing = HEADER_COMPLETE;
			m->size_remaining = load_32(ib + 4);
			m->status_eof = m->size_remaining == 0 ? LZO_EOF_YES
							       : LZO_EOF_NO;
			m->input_buffer_filled = 0;
			break;
		default:
			RAISE(msg, LZO_CONSTRAINT_ERROR,
			      "Header_Complete case must not be reached while "
			      "processing header bytes. Library bug detected.");
		}
	}
	return LZO_OK;
}

/* ---------------------------------------------------------------- Init -- */

static void meta_defaults(struct meta *m, int reservation)
{
	memset(m, 0, sizeof *m);
	m->is_format = F_TBD;
	m->header_parsing = NEED_MAGIC;
	m->memory_reservation = reservation;
	m->status_eof = LZO_EOF_NO;
	m->size_remaining = 4;
}

static lzo_ctx *new_ctx(const struct meta *m, int in_last)
{
	lzo_ctx *c = (lzo_ctx *)calloc(1, sizeof *c);
	if (!c)
		return NULL;
	c->m = *m;
	c->input_buffer_len = in_last + 1;
	c->input_buffer = (uint8_t *)calloc(1, (size_t)c->input_buffer_len + 32);
	if (!c->input_buffer) {
		free(c);
		return NULL;
	}
	c->input_length = -1;
	lzo_xxh32_reset(&c->hash_all_data, 0);
	return c;
}

lzo_ctx *lzo_init(int reservation, int *min_buffer_size)   /* :48-63 */
{
	struct meta m;
	int block_max = get_block_size(reservation);
	meta_defaults(&m, reservation);
	*min_buffer_size = block_max + HISTORY_SIZE + 8;
	return new_ctx(&m, block_max + 4 + BLOCK_SIZE_BYTES - 1);
}

lzo_ctx *lzo_init_with_header(const uint8_t *input, int input_len,
		int *num_consumed, int *min_buffer_size, int reservation,
		int *exc, char *msg, size_t msg_cap)   /* :79-125 */
{
	uint8_t header_buffer[20];
	char lmsg[640];
	struct meta mt;
	int pos = 0, inner = 0, block_max, rc;

	lmsg[0] = 0;
	*exc = LZO_OK;
	*num_consumed = 0;
	if (input_len < 7) {   /* Pre => Input'Length >= 7, lib/lz4ada.ads:243 */
		*exc = raise_to(lmsg, sizeof lmsg, LZO_ASSERTION_ERROR,
				"failed precondition from lz4ada.ads:243");
		goto fail;
	}
	meta_defaults(&mt, reservation == LZO_SINGLE_FRAME ? LZO_USE_FIRST
							    : reservation);
	while (mt.header_parsing != HEADER_COMPLETE) {
		if (pos >= input_len) {
			*exc = raise_to(lmsg, sizeof lmsg,
				LZO_TOO_FEW_HEADER_BYTES,
				"Expected at least  %llu more bytes but header "
				"input has already ended.",
				(unsigned long long)mt.size_remaining);
			goto fail;
		}
		rc = process_header_bytes(&mt, lmsg, header_buffer, input + pos,
					  input_len - pos, &inner);
		if (rc != LZO_OK) {
			*exc = rc;
			goto fail;
		}
		pos += inner;
		*num_consumed += inner;
	}
	block_max = get_block_size(mt.memory_reservation);
	*min_buffer_size = block_max + HISTORY_SIZE + 8;
	{
		int in_last = block_max + mt.block_checksum_length +
			      BLOCK_SIZE_BYTES - 1;
		if (reservation == LZO_SINGLE_FRAME)
			mt.memory_reservation = LZO_SINGLE_FRAME;
		return new_ctx(&mt, in_last);
	}
fail:
	if (msg && msg_cap) {
		strncpy(msg, lmsg, msg_cap - 1);
		msg[msg_cap - 1] = 0;
	}
	return NULL;
}

lzo_ctx *lzo_init_for_block(int *min_buffer_size, int compressed_length,
		int reservation)   /* :127-147 */
{
	struct meta m;
	lzo_ctx *c;
	int block_max = get_block_size(reservation);
	meta_defaults(&m, reservation);
	m.is_format = F_BLOCK;
	m.is_compressed = 1;
	m.header_parsing = HEADER_COMPLETE;
	*min_buffer_size = block_max + HISTORY_SIZE + 8;
	c = new_ctx(&m, block_max - 1);
	if (c)
		c->input_length = compressed_length;
	return c;
}

void lzo_free(lzo_ctx *c)
{
	if (c) {
		free(c->input_buffer);
		free(c);
	}
}

const char *lzo_message(const lzo_ctx *c) { return c->msg; }

/* -------------------------------------------------------- Block decode -- */

/* Decrease_Data_Size_Remaining, lib/lz4ada.adb:826-839 */
static int decrease_data_size_remaining(lzo_ctx *c, uint64_t n)
{
	if (c->m.has_content_size) {
		if (c->m.size_remaining < n)
			RAISE(c->msg, LZO_DATA_CORRUPTION,
			      "Produced content size exceeds declared content "
			      "size. The supplied data is inconsistent.");
		c->m.size_remaining -= n;
	}
	return LZO_OK;
}

/* The reference copies 8-byte slices (lib/lz4ada.adb:811-817) and may over-write up to seven
 * bytes; this copy moves 8-byte slices too but never writes outside [dst, dst + n): the last
 * slice is placed so that it ends exactly at the end.  Ranges that overlap closer than eight
 * bytes (history replay with src just ahead of dst) fall back to memmove. */
static inline void copy_exact(uint8_t *dst, const uint8_t *src, int n)
{
	ptrdiff_t gap = src > dst ? src - dst : dst - src;
	if (gap < 8 || n > 64) {
		memmove(dst, src, (size_t)n);
	} else if (n >= 8) {
		int i;
		uint64_t last, v;
		memcpy(&last, src + n - 8, 8);
		for (i = 0; i + 8 < n; i += 8) {
			memcpy(&v, src + i, 8);
			memcpy(dst + i, &v, 8);
		}
		memcpy(dst + n - 8, &last, 8);
	} else if (n >= 4) {           /* two overlapping 4-byte moves */
		uint32_t a, b;
		memcpy(&a, src, 4);
		memcpy(&b, src + n - 4, 4);
		memcpy(dst, &a, 4);
		memcpy(dst + n - 4, &b, 4);
	} else if (n >= 2) {
		uint16_t a, b;
		memcpy(&a, src, 2);
		memcpy(&b, src + n - 2, 2);
		memcpy(dst, &a, 2);
		memcpy(dst + n - 2, &b, 2);
	} else if (n == 1) {
		dst[0] = src[0];
	}
}

/* Write_Output, lib/lz4ada.adb:790-824 -- exact copy (no 8-byte over-copy),
 * content-size accounting first, then a capacity check the reference lacks
 * (Appendix C). */
static int write_output(lzo_ctx *c, const uint8_t *src, int n,
			uint8_t *buffer, int buffer_len)
{
	TRY(decrease_data_size_remaining(c, (uint64_t)n));
	if ((int64_t)c->output_pos + n > buffer_len)
		RAISE(c->msg, LZO_DATA_CORRUPTION,
		      "Output buffer exhausted. Decompressed data does not fit "
		      "into the %d bytes provided.", buffer_len);
	copy_exact(buffer + c->output_pos, src, n);
	c->output_pos += n;
	return LZO_OK;
}

/* Output_With_History, lib/lz4ada.adb:845-904 */
static int output_with_history(lzo_ctx *c, int offset, int match_length,
			       uint8_t *buffer, int buffer_len)
{
	int raw_offset = c->output_pos - offset;
	int remaining = match_length;
	int i_offset, i_length;

	if (raw_offset >= 0) {
		i_offset = raw_offset;
		i_length = match_length < offset ? match_length : offset;
	} else {
		int h_offset = raw_offset + c->output_pos_history;
		int h_length = offset - c->output_pos;
		if (match_length < h_length)
			h_length = match_length;
		if (h_offset < 0)
			RAISE(c->msg, LZO_DATA_CORRUPTION,
			      "Backreference location out of range. Read from "
			      "offset %d not possible (earliest available "
			      "index is 0).", h_offset);
		if (h_length > 0) {
			TRY(write_output(c, buffer + h_offset, h_length, buffer,
					 buffer_len));
			remaining = match_length - h_length;
		}
		i_offset = 0;
		i_length = remaining < c->output_pos ? remaining
						     : c->output_pos;
	}
	if (i_length > 0) {
		TRY(write_output(c, buffer + i_offset, i_length, buffer,
				 buffer_len));
		remaining -= i_length;
	}
	if (remaining > 0) {   /* repeating part, :893-903 */
		int r_start = c->output_pos - offset;
		int processed = 0;
		while (processed < remaining) {
			int r_len = c->output_pos - r_start;
			if (remaining - processed < r_len)
				r_len = remaining - processed;
			TRY(write_output(c, buffer + r_start, r_len, buffer,
					 buffer_len));
			processed += r_len;
		}
	}
	return LZO_OK;
}

/* Update_Checksum, lib/lz4ada.adb:709-714 */
static void update_checksum(lzo_ctx *c, const uint8_t *p, int n)
{
	if (c->m.content_checksum_length != 0 && n > 0)
		lzo_xxh32_update(&c->hash_all_data, p, (size_t)n);
}

/* Process_Variable_Length, lib/lz4ada.adb:724-735, bounds-checked */
static int process_variable_length(lzo_ctx *c, const uint8_t *raw, int n,
				   int *idx, int64_t *var, const char *what)
{
	if (*var == 15) {
		uint8_t tmp;
		do {
			if (*idx >= n)
				RAISE(c->msg, LZO_DATA_CORRUPTION,
				      "Corrupted Block: %s length extension "
				      "runs past the end of the block.", what);
			tmp = raw[*idx];
			*var += tmp;
			*idx += 1;
		} while (tmp == 255);
	}
	return LZO_OK;
}

/* Decompress_Full_Block + Decompress_Sequence, lib/lz4ada.adb:716-788 */
static int decompress_full_block(lzo_ctx *c, const uint8_t *raw, int n,
				 uint8_t *buffer, int buffer_len,
				 int *output_first, int *output_last)
{
	int idx = 0;
	*output_first = c->output_pos;
	while (idx < n) {
		uint8_t token = raw[idx];
		int64_t num_literals = (token & 0xf0) >> 4;
		int64_t match_length = token & 0x0f;
		int raw_nibble = token & 0x0f;
		int offset;
		idx += 1;
		TRY(process_variable_length(c, raw, n, &idx, &num_literals,
					    "Literal"));
		if (num_literals > 0) {
			if (num_literals > (int64_t)(n - idx)) {
				/* The reference reads past the block here
				 * (checks suppressed, :798-801), still does
				 * the content-size accounting, and then
				 * reports :754 when the nibble is non-zero. */
				TRY(decrease_data_size_remaining(c,
						(uint64_t)num_literals));
				if (raw_nibble != 0)
					goto ends_after_literals;
				RAISE(c->msg, LZO_DATA_CORRUPTION,
				      "Corrupted Block: Literal run of %lld "
				      "bytes exceeds the %d bytes left in the "
				      "block.", (long long)num_literals,
				      n - idx);
			}
			TRY(write_output(c, raw + idx, (int)num_literals, buffer,
					 buffer_len));
			idx += (int)num_literals;
		}
		if (idx >= n) {   /* :752-764 */
			if (raw_nibble != 0) {
ends_after_literals:
				RAISE(c->msg, LZO_DATA_CORRUPTION,
				      "Match_Length= %d suggests compressed "
				      "data but this sequence already ends "
				      "after the literals. This might also "
				      "happen with an untypical encoder?",
				      raw_nibble);
			}
			break;
		}
		if (idx + 1 >= n)   /* reference: Constraint_Error, :766-767 */
			RAISE(c->msg, LZO_DATA_CORRUPTION,
			      "Corrupted Block: Block ends inside a match "
			      "offset.");
		offset = (int)raw[idx] | ((int)raw[idx + 1] << 8);
		idx += 2;
		if (offset == 0)
			RAISE(c->msg, LZO_DATA_CORRUPTION,
			      "Corrupted Block: Offset = 0 detected.");
		TRY(process_variable_length(c, raw, n, &idx, &match_length,
					    "Match"));
		TRY(output_with_history(c, offset, (int)match_length + 4, buffer,
					buffer_len));
	}
	*output_last = c->output_pos - 1;
	update_checksum(c, buffer + *output_first,
			*output_last - *output_first + 1);
	if (c->output_pos >= HISTORY_SIZE)
		c->output_pos_history = c->output_pos;
	return LZO_OK;
}

/* Decode_Full_Block_With_Trailer, lib/lz4ada.adb:661-696 with
 * Check_Checksum (:698-707) folded in */
static int decode_full_block_with_trailer(lzo_ctx *c, const uint8_t *blk,
		int blk_len, uint8_t *buffer, int buffer_len,
		int *output_first, int *output_last)
{
	int raw_len = blk_len - c->m.block_checksum_length;
	if (c->m.block_checksum_length > 0) {
		uint32_t expect = load_32(blk + raw_len);
		uint32_t computed = lzo_xxh32_hash(blk, (size_t)raw_len);
		if (computed != expect)
			RAISE(c->msg, LZO_CHECKSUM_ERROR,
			      "Declared checksum is 0x%08x, but computed one "
			      "is 0x%08x.", expect, computed);
	}
	if (c->output_pos >= HISTORY_SIZE)
		c->output_pos = 0;
	if (c->m.is_compressed) {
		TRY(decompress_full_block(c, blk, raw_len, buffer, buffer_len,
					  output_first, output_last));
	} else {
		TRY(write_output(c, blk, raw_len, buffer, buffer_len));
		if (c->output_pos >= HISTORY_SIZE)
			c->output_pos_history = c->output_pos;
		*output_first = c->output_pos - raw_len;
		*output_last = c->output_pos - 1;
		update_checksum(c, buffer + *output_first, raw_len);
	}
	return LZO_OK;
}

/* -------------------------------------------------------------- Update -- */

static int reset_for_next_frame(lzo_ctx *c, const uint8_t *input,
				int input_len, int *num_consumed);

/* Reset_Outer_For_Next_Frame, lib/lz4ada.adb:451-461 */
static void reset_outer_for_next_frame(lzo_ctx *c)
{
	c->is_at_end_mark = 0;
	c->input_length = -1;
	c->output_pos = 0;
	c->output_pos_history = 0;
	lzo_xxh32_reset(&c->hash_all_data, 0);
}

/* Reset_For_Next_Frame, lib/lz4ada.adb:435-449 */
static int reset_for_next_frame(lzo_ctx *c, const uint8_t *input,
				int input_len, int *num_consumed)
{
	if (c->m.memory_reservation == LZO_SINGLE_FRAME)
		RAISE(c->msg, LZO_DATA_CORRUPTION,
		      "Requested Single_Frame operation but data was provided "
		      "after End of Frame was detected");
	c->m.status_eof = LZO_EOF_NO;
	c->m.header_parsing = NEED_MAGIC;
	c->m.size_remaining = 4;
	reset_outer_for_next_frame(c);
	return process_header_bytes(&c->m, c->msg, c->input_buffer, input,
				    input_len, num_consumed);
}

/* Skip, lib/lz4ada.adb:420-433 */
static int skip(lzo_ctx *c, const uint8_t *input, int input_len,
		int *num_consumed)
{
	uint64_t remain = c->m.size_remaining;
	uint64_t consumed = (uint64_t)input_len < remain ? (uint64_t)input_len
							 : remain;
	if (c->m.status_eof == LZO_EOF_YES && consumed == 0)
		return reset_for_next_frame(c, input, input_len, num_consumed);
	*num_consumed = (int)consumed;
	c->m.size_remaining = remain - consumed;
	c->m.status_eof = c->m.size_remaining == 0 ? LZO_EOF_YES : LZO_EOF_NO;
	return LZO_OK;
}

/* Check_End_Mark, lib/lz4ada.adb:463-523 */
static int set_frame_has_ended(lzo_ctx *c)
{
	c->m.status_eof = LZO_EOF_YES;
	c->m.input_buffer_filled = 0;
	if (c->m.has_content_size && c->m.size_remaining != 0)
		RAISE(c->msg, LZO_DATA_CORRUPTION,
		      "Frame has ended, but according to content size, there "
		      "should be  %llu bytes left to output.",
		      (unsigned long long)c->m.size_remaining);
	return LZO_OK;
}

static int check_end_mark(lzo_ctx *c, const uint8_t *input, int input_len,
			  int *num_consumed)
{
	int provided = input_len - *num_consumed;
	int required = c->m.content_checksum_length - c->m.input_buffer_filled;

	if (c->m.content_checksum_length == 0 ||
	    c->m.status_eof == LZO_EOF_YES || required <= 0) {
		if (c->m.status_eof == LZO_EOF_YES) {
			if (*num_consumed != 0)
				RAISE(c->msg, LZO_ASSERTION_ERROR,
				      "Num_Consumed = 0");
			return reset_for_next_frame(c, input, input_len,
						    num_consumed);
		}
		return set_frame_has_ended(c);
	} else if (provided >= required) {
		uint8_t word[4];
		uint32_t checksum, compare;
		memcpy(word, c->input_buffer, (size_t)c->m.input_buffer_filled);
		memcpy(word + c->m.input_buffer_filled, input + *num_consumed,
		       (size_t)required);
		checksum = load_32(word);
		compare = lzo_xxh32_final(&c->hash_all_data);
		*num_consumed += required;
		if (checksum != compare)
			RAISE(c->msg, LZO_CHECKSUM_ERROR,
			      "Computed content checksum 0x%08x does not match "
			      "declared content checksum 0x%08x.", compare,
			      checksum);
		return set_frame_has_ended(c);
	}
	memcpy(c->input_buffer + c->m.input_buffer_filled,
	       input + *num_consumed, (size_t)provided);
	c->m.input_buffer_filled += provided;
	*num_consumed += provided;
	return LZO_OK;
}

/* Is_Any_Magic_Number, lib/lz4ada.adb:587-593 */
static int is_any_magic_number(uint32_t v)
{
	return v == MAGIC_MODERN || v == MAGIC_LEGACY ||
	       (v >= MAGIC_SKIP_LO && v <= MAGIC_SKIP_HI);
}

/* Try_Detect_Input_Length, lib/lz4ada.adb:525-585 */
static int try_detect_input_length(lzo_ctx *c, const uint8_t *input,
				   int input_len, int *num_consumed)
{
	int additional = BLOCK_SIZE_BYTES + c->m.block_checksum_length;
	uint32_t length_word;
	int take = BLOCK_SIZE_BYTES - c->m.input_buffer_filled;
	if (input_len < take)
		take = input_len;
	*num_consumed = take;
	memcpy(c->input_buffer + c->m.input_buffer_filled, input, (size_t)take);
	c->m.input_buffer_filled += take;

	if (c->m.input_buffer_filled != BLOCK_SIZE_BYTES)
		return LZO_OK;
	length_word = load_32(c->input_buffer);
	if (c->m.is_format == F_MODERN && length_word == 0) {
		c->is_at_end_mark = 1;
		c->m.input_buffer_filled = 0;
	} else if (c->m.is_format == F_LEGACY &&
		   is_any_magic_number(length_word)) {
		if (c->m.memory_reservation == LZO_SINGLE_FRAME)
			RAISE(c->msg, LZO_DATA_CORRUPTION,
			      "Requested Single_Frame operation but data "
			      "provided what looks like the beginning of "
			      "another frame.");
		reset_outer_for_next_frame(c);
		TRY(process_header_magic(&c->m, c->msg, length_word));
	} else {   /* Detect_Modern, :531-554 */
		if (c->m.is_format == F_MODERN) {
			c->m.is_compressed = (length_word & 0x80000000u) == 0;
			length_word &= 0x7ffffffu;   /* 27 bits, :538 */
		}
		if (length_word > 0x7fffffffu)   /* reference: Constraint_Error */
			RAISE(c->msg, LZO_DATA_CORRUPTION,
			      "Declared block length %u is not representable.",
			      length_word);
		c->input_length = (int)length_word;
		if ((int64_t)c->input_length + additional >
		    c->input_buffer_len) {
			c->input_length = -1;
			RAISE(c->msg, LZO_DATA_CORRUPTION,
			      "Declared maximum data length exceeded. Buffer "
			      "has  %d bytes, current block requires  %u bytes "
			      "+  %d bytes for metadata.",
			      c->input_buffer_len, length_word, additional);
		}
	}
	return LZO_OK;
}

/* Cache_Data_And_Process_If_Full, lib/lz4ada.adb:630-659 */
static int cache_data_and_process_if_full(lzo_ctx *c, const uint8_t *input,
		int input_len, int *num_consumed, uint8_t *buffer,
		int buffer_len, int *output_first, int *output_last)
{
	int avail = input_len - *num_consumed;
	int head = c->m.is_format == F_BLOCK ? 0 : BLOCK_SIZE_BYTES;
	int want = c->input_length + c->m.block_checksum_length -
		   c->m.input_buffer_filled + head;
	int fill = c->m.input_buffer_filled;
	const uint8_t *src = input + *num_consumed;

	/* Init_For_Block never checks Compressed_Length against In_Last; the
	 * reference would fail with Constraint_Error on the slice below. */
	if ((int64_t)fill + (want > avail ? avail : want) > c->input_buffer_len)
		RAISE(c->msg, LZO_DATA_CORRUPTION,
		      "Declared block length exceeds the input buffer of  %d "
		      "bytes.", c->input_buffer_len);
	if (want > avail) {
		memcpy(c->input_buffer + fill, src, (size_t)avail);
		c->m.input_buffer_filled += avail;
		*num_consumed += avail;
		return LZO_OK;
	}
	*num_consumed += want;
	c->m.input_buffer_filled = 0;
	c->input_length = -1;
	/* Input_Buffer(head .. Fill-1) & Input(Offset .. Offset+Want-1); the
	 * input buffer has room for the whole block, so append in place. */
	memcpy(c->input_buffer + fill, src, (size_t)want);
	return decode_full_block_with_trailer(c, c->input_buffer + head,
			fill - head + want, buffer, buffer_len, output_first,
			output_last);
}

/* Handle_Newly_Known_Input_Length, lib/lz4ada.adb:595-628 */
static int handle_newly_known_input_length(lzo_ctx *c, const uint8_t *input,
		int input_len, int *num_consumed, uint8_t *buffer,
		int buffer_len, int *output_first, int *output_last)
{
	int total = c->input_length + c->m.block_checksum_length;
	if (input_len - *num_consumed >= total) {
		const uint8_t *use = input + *num_consumed;
		*num_consumed += total;
		c->m.input_buffer_filled = 0;
		c->input_length = -1;
		return decode_full_block_with_trailer(c, use, total, buffer,
				buffer_len, output_first, output_last);
	}
	return cache_data_and_process_if_full(c, input, input_len,
			num_consumed, buffer, buffer_len, output_first,
			output_last);
}

/* Update, lib/lz4ada.adb:383-418 */
int lzo_update(lzo_ctx *c, const uint8_t *input, int input_len,
	       int *num_consumed, uint8_t *buffer, int buffer_len,
	       int *output_first, int *output_last)
{
	*num_consumed = 0;
	*output_first = 1;
	*output_last = 0;
	c->msg[0] = 0;
	if (c->m.header_parsing != HEADER_COMPLETE)
		return process_header_bytes(&c->m, c->msg, c->input_buffer,
					    input, input_len, num_consumed);
	if (c->m.is_format == F_SKIPPABLE)
		return skip(c, input, input_len, num_consumed);
	if (c->m.is_format == F_TBD)
		RAISE(c->msg, LZO_ASSERTION_ERROR, "Is_Format /= TBD");
	if (c->is_at_end_mark)
		return check_end_mark(c, input, input_len, num_consumed);
	if (c->input_length != -1)
		return cache_data_and_process_if_full(c, input, input_len,
				num_consumed, buffer, buffer_len, output_first,
				output_last);
	TRY(try_detect_input_length(c, input, input_len, num_consumed));
	if (c->is_at_end_mark)
		return check_end_mark(c, input, input_len, num_consumed);
	if (c->input_length != -1)
		return handle_newly_known_input_length(c, input, input_len,
				num_consumed, buffer, buffer_len, output_first,
				output_last);
	return LZO_OK;
}

/* Is_End_Of_Frame, lib/lz4ada.adb:906-915 */
int lzo_is_end_of_frame(const lzo_ctx *c)
{
	switch (c->m.is_format) {
	case F_LEGACY:
		return c->is_at_end_mark ? LZO_EOF_MAYBE : c->m.status_eof;
	case F_BLOCK:
		return c->input_length == -1 ? LZO_EOF_YES : LZO_EOF_NO;
	default:
		return c->m.status_eof;
	}
}

/* ------------------------------------------------- test-flow drivers -- */

static void copy_msg(char *dst, size_t cap, const char *src)
{
	if (dst && cap) {
		strncpy(dst, src, cap - 1);
		dst[cap - 1] = 0;
	}
}

static int drive(lzo_ctx *c, const uint8_t *in, size_t in_len, size_t chunk,
		 uint8_t *ring, int ring_len, uint8_t *out, size_t out_cap,
		 size_t *out_len, int require_progress)
{
	size_t pos = 0;
	int idle = 0;
	while (pos < in_len) {
		size_t end = pos + chunk;
		if (end > in_len || end < pos)
			end = in_len;
		while (pos < end) {
			int consumed, of, ol, rc, n;
			rc = lzo_update(c, in + pos, (int)(end - pos), &consumed,
					ring, ring_len, &of, &ol);
			if (rc != LZO_OK)
				return rc;
			n = ol - of + 1;
			if (n > 0) {
				if (*out_len + (size_t)n > out_cap)
					return raise_to(c->msg, sizeof c->msg,
						LZO_CONSTRAINT_ERROR,
						"oracle driver: output "
						"capacity exceeded");
				memcpy(out + *out_len, ring + of, (size_t)n);
				*out_len += (size_t)n;
			}
			pos += (size_t)consumed;
			if (consumed == 0 && n <= 0) {
				if (require_progress || ++idle > 4)
					return raise_to(c->msg, sizeof c->msg,
						LZO_ASSERTION_ERROR,
						"No more data accepted but no "
						"exception signalled.");
			} else {
				idle = 0;
			}
		}
	}
	return LZO_OK;
}

int lzo_decode_stream(const uint8_t *in, size_t in_len, size_t chunk,
		uint8_t *out, size_t out_cap, size_t *out_len, int *eof,
		char *msg, size_t msg_cap)
{
	int min_buf, rc;
	lzo_ctx *c = lzo_init(LZO_FOR_ALL, &min_buf);
	uint8_t *ring;
	if (!c)
		return LZO_CONSTRAINT_ERROR;
	ring = (uint8_t *)malloc((size_t)min_buf);
	*out_len = 0;
	if (chunk == 0)
		chunk = in_len ? in_len : 1;
	rc = drive(c, in, in_len, chunk, ring, min_buf, out, out_cap, out_len,
		   0);
	if (eof)
		*eof = lzo_is_end_of_frame(c);
	copy_msg(msg, msg_cap, c->msg);
	free(ring);
	lzo_free(c);
	return rc;
}

int lzo_decode_error_case(const uint8_t *in, size_t in_len,
		uint8_t *out, size_t out_cap, size_t *out_len,
		char *msg, size_t msg_cap)
{
	int consumed = 0, min_buf = 0, exc = LZO_OK, rc;
	uint8_t *ring;
	lzo_ctx *c = lzo_init_with_header(in, (int)in_len, &consumed, &min_buf,
					  LZO_SINGLE_FRAME, &exc, msg, msg_cap);
	*out_len = 0;
	if (!c)
		return exc;
	ring = (uint8_t *)malloc((size_t)min_buf);
	rc = drive(c, in + consumed, in_len - (size_t)consumed,
		   in_len - (size_t)consumed, ring, min_buf, out, out_cap,
		   out_len, 1);
	copy_msg(msg, msg_cap, c->msg);
	free(ring);
	lzo_free(c);
	return rc;
}
