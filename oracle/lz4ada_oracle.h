/*
 * lz4ada_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the observable behaviour of the reference Ada
 * package LZ4Ada (/root/reference/lib/lz4ada.ads, lib/lz4ada.adb).  It exists
 * so that the CUDA path can be checked bit-for-bit and message-for-message
 * against "what the reference would have done".  Only tests/, the smoke check
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (liblz4b200.so) never links or calls it.
 *
 * Parity pinning: the reference is 100 % Ada and there is no GNAT in this
 * image, so the reference itself cannot be executed here.  The oracle is
 * pinned instead against every golden vector the reference's own test-suite
 * holds (24 .lz4/.bin pairs under 4 KiB and 1-byte feeding, 15 .err/.eds
 * pairs with exact messages, the XXH32 KAT, the inline two-legacy-frame and
 * raw-block cases of test_suite/lz4test.adb) -- see tests/test_oracle_*.py.
 *
 * Deliberate deviations (SURVEY.md Appendix C -- reference defects that the
 * reference's own tests never exercise; "parity unpinned" for these only):
 *   - exact copies instead of the 8-byte over-copy (lib/lz4ada.adb:811-817),
 *     which removes the history clobber for distances 65530..65535;
 *   - every read of the compressed block and every write of the output is
 *     bounds-checked and reported as Data_Corruption with a new message;
 *   - raw-block API fed in several chunks keeps its first four cached bytes
 *     (lib/lz4ada.adb:654 drops them);
 *   - legacy length words >= 2**31 are Data_Corruption, not Constraint_Error.
 */
#ifndef LZ4ADA_ORACLE_H
#define LZ4ADA_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/lz4ada.ads:79-80 (same order, same 'Image spellings) */
enum lzo_reservation {
	LZO_SZ_64_KIB = 0, LZO_SZ_256_KIB, LZO_SZ_1_MIB, LZO_SZ_4_MIB,
	LZO_SZ_8_MIB, LZO_USE_FIRST, LZO_SINGLE_FRAME
};
#define LZO_FOR_MODERN LZO_SZ_4_MIB   /* lib/lz4ada.ads:92  */
#define LZO_FOR_LEGACY LZO_SZ_8_MIB   /* lib/lz4ada.ads:100 */
#define LZO_FOR_ALL    LZO_SZ_8_MIB   /* lib/lz4ada.ads:106 */

/* lib/lz4ada.ads:124 */
enum lzo_end_of_frame { LZO_EOF_YES = 0, LZO_EOF_NO = 1, LZO_EOF_MAYBE = 2 };

/* lib/lz4ada.ads:133-162 plus the two non-library outcomes a caller can hit */
enum lzo_exception {
	LZO_OK = 0,
	LZO_CHECKSUM_ERROR,
	LZO_DATA_CORRUPTION,
	LZO_NOT_SUPPORTED,
	LZO_TOO_FEW_HEADER_BYTES,
	LZO_TOO_LITTLE_MEMORY,
	LZO_CONSTRAINT_ERROR,   /* "Library bug detected", lib/lz4ada.adb:185 */
	LZO_ASSERTION_ERROR     /* violated Pre / Ada.Assertions.Assert */
};

typedef struct lzo_ctx lzo_ctx;

/* Init, lib/lz4ada.adb:48-63 */
lzo_ctx *lzo_init(int reservation, int *min_buffer_size);

/* Init_With_Header, lib/lz4ada.adb:79-125.  On failure returns NULL and
 * fills exc (enum lzo_exception) and msg ("raised LZ4ADA.X : text"). */
lzo_ctx *lzo_init_with_header(const uint8_t *input, int input_len,
		int *num_consumed, int *min_buffer_size, int reservation,
		int *exc, char *msg, size_t msg_cap);

/* Init_For_Block, lib/lz4ada.adb:127-147 */
lzo_ctx *lzo_init_for_block(int *min_buffer_size, int compressed_length,
		int reservation);

/* Update (Octets flavour), lib/lz4ada.adb:383-418.  Buffer'First = 0.
 * Returns enum lzo_exception; on error the message is lzo_message(ctx). */
int lzo_update(lzo_ctx *ctx, const uint8_t *input, int input_len,
		int *num_consumed, uint8_t *buffer, int buffer_len,
		int *output_first, int *output_last);

/* Is_End_Of_Frame, lib/lz4ada.adb:906-915 */
int lzo_is_end_of_frame(const lzo_ctx *ctx);

/* "raised LZ4ADA.<NAME> : <message>" exactly as GNAT's Exception_Information
 * first line (test_suite/lz4test.adb:310-323 strips the trailing LF). */
const char *lzo_message(const lzo_ctx *ctx);

void lzo_free(lzo_ctx *ctx);

/* XXHash32, lib/lz4ada.adb:923-1026 */
typedef struct {
	uint32_t state[4];
	uint8_t  buffer[16];
	int      buffer_size;
	uint64_t total_length;
} lzo_xxh32;
void     lzo_xxh32_reset(lzo_xxh32 *h, uint32_t seed);         /* :932 */
void     lzo_xxh32_update(lzo_xxh32 *h, const uint8_t *p, size_t n); /* :942 */
uint32_t lzo_xxh32_final(const lzo_xxh32 *h);                  /* :993 */
uint32_t lzo_xxh32_hash(const uint8_t *p, size_t n);           /* :1019 */

/* Convenience for tests and the CPU baseline: run a whole stream the way
 * tool_unlz4ada_simple/unlz4ada_simple.adb:23-36 and
 * test_suite/lz4test.adb:32-83 do (Init(For_All) + Update loop fed `chunk`
 * bytes at a time), appending every output slice to out.  Returns enum
 * lzo_exception; *out_len is the number of bytes produced before any error;
 * *eof is Is_End_Of_Frame after the last call; msg receives the exception
 * text. */
int lzo_decode_stream(const uint8_t *in, size_t in_len, size_t chunk,
		uint8_t *out, size_t out_cap, size_t *out_len, int *eof,
		char *msg, size_t msg_cap);

/* The error-vector flow of test_suite/lz4test.adb:280-308:
 * Init_With_Header(all, Single_Frame) then Update on the rest. */
int lzo_decode_error_case(const uint8_t *in, size_t in_len,
		uint8_t *out, size_t out_cap, size_t *out_len,
		char *msg, size_t msg_cap);

#ifdef __cplusplus
}
#endif
#endif
