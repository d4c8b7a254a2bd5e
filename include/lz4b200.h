/*
 * lz4b200.h -- C-ABI of the B200-native LZ4 decompressor (liblz4b200.so).
 *
 * Two layers, both plain C (pointers and sizes only, no C++/torch types, no
 * callbacks, no exceptions across the boundary):
 *
 *   lz4b200_*  the device shim.  This is what the Ada host layer binds with
 *              `pragma Import (C, ...)`: context, device
 *              buffers, transfers and the kernel launches K1..K5.  It has no
 *              counterpart in the reference (which is CPU-only); each entry
 *              names the reference code whose work it takes over.
 *   lz4ada_*   the LZ4Ada package API (reference lib/lz4ada.ads:50-344) with
 *              the same names, argument meaning and error behaviour, written
 *              in C++ above the shim because this image has no Ada compiler.
 *              Ada exceptions become a return code (enum lz4ada_exception)
 *              plus the exact GNAT Exception_Information line.
 *
 * There is no CPU decode path behind any of these calls: block payloads are
 * only ever decoded by the sm_100a kernels.  If the CUDA runtime or a device
 * is missing every call fails with LZ4B200_ERR_CUDA.
 */
#ifndef LZ4B200_H
#define LZ4B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZ4B200_ABI_VERSION 1

/* ------------------------------------------------------------------------
 * Device shim
 * --------------------------------------------------------------------- */

/* Return codes of every lz4b200_* call.  LZ4 *data* errors never travel here;
 * they are reported per block in lz4b200_blk_status. */
enum lz4b200_rc {
	LZ4B200_OK          = 0,
	LZ4B200_ERR_CUDA    = -1,   /* CUDA runtime/driver failure; see lz4b200_last_error */
	LZ4B200_ERR_ARG     = -2,
	LZ4B200_ERR_NOMEM   = -3
};

typedef struct lz4b200_ctx lz4b200_ctx;

/* One descriptor per LZ4 block, built by the host block-table builder from the
 * 4-byte size words the reference reads in Try_Detect_Input_Length
 * (lib/lz4ada.adb:525-585).  Offsets are relative to the src / dst base
 * pointers given to the launch. */
typedef struct lz4b200_blk_desc {
	uint64_t src_off;     /* first payload byte (after the size word) */
	uint64_t dst_off;     /* where this block's output starts */
	uint32_t src_len;     /* payload bytes, excluding the optional checksum trailer */
	uint32_t dst_cap;     /* output bytes this block may produce */
	uint32_t flags;       /* LZ4B200_BLK_* */
	uint32_t hist_avail;  /* bytes of this frame's output that precede dst_off and may be
	                       * referenced (frame position of the block start; 0xffffffff once
	                       * the reference's 64 KiB ring has wrapped, lib/lz4ada.adb:864) */
} lz4b200_blk_desc;

#define LZ4B200_BLK_STORED        1u  /* size word bit 31 set, lib/lz4ada.adb:536-537 */
#define LZ4B200_BLK_HAS_CHECKSUM  2u  /* 4-byte LE XXH32 trailer follows the payload, :672-676 */
#define LZ4B200_BLK_HASH_ONLY     4u  /* verify checksum, do not decode (used by retries) */
#define LZ4B200_BLK_CHAINED       8u  /* decoded in order by lz4b200_decode_linked; K1 skips it */
#define LZ4B200_BLK_FIRST_OF_FRAME 16u /* chain kernel: a new frame starts here, history restarts */
#define LZ4B200_BLK_RING_CAP      64u /* chain kernel: dst_cap is the length of the reference caller's Buffer
                                       * (Min_Buffer_Size = reservation block size + 64 KiB + 8, lib/lz4ada.adb:54);
                                       * the block may produce what is left of it behind the ring cursor
                                       * (:678-680) -- the only bound the reference puts on a block's output */
#define LZ4B200_BLK_K2           128u /* a stored block the host has routed to lz4b200_copy_stored; K1 skips it */
#define LZ4B200_BLK_NOT_K1 (LZ4B200_BLK_CHAINED | LZ4B200_BLK_K2)
#define LZ4B200_BLK_SOLO          32u /* chain of one block taken from an independent frame (big blocks get a
                                       * whole CTA): its exact path keeps independent-block semantics */

/* Per-block outcome written by the kernels; the host folds these in stream
 * order into the reference's exceptions. */
typedef struct lz4b200_blk_status {
	uint32_t code;            /* LZ4B200_ST_* */
	uint32_t out_len;         /* bytes produced (valid when code == OK) */
	uint32_t err_pos;         /* block-relative output position when the error was detected
	                           * (after the literals of the failing sequence) */
	int32_t  aux;             /* value the reference prints: raw match nibble (:754),
	                           * frame_pos - offset (:868) */
	uint32_t xxh32_computed;  /* block checksum as computed (when HAS_CHECKSUM) */
	uint32_t xxh32_declared;  /* block checksum as stored in the trailer */
} lz4b200_blk_status;

enum lz4b200_status_code {
	LZ4B200_ST_OK                  = 0,
	LZ4B200_ST_BLOCK_CHECKSUM      = 1,  /* lib/lz4ada.adb:702 */
	LZ4B200_ST_ENDS_AFTER_LITERALS = 2,  /* :754 */
	LZ4B200_ST_OFFSET_ZERO         = 3,  /* :770 */
	LZ4B200_ST_BACKREF_RANGE       = 4,  /* :868 */
	LZ4B200_ST_LITERAL_OVERRUN     = 5,  /* literal run past block end with nibble 0 (ref: silent garbage) */
	LZ4B200_ST_LIT_EXT_OVERRUN     = 6,  /* length extension past block end (ref: Constraint_Error) */
	LZ4B200_ST_MATCH_EXT_OVERRUN   = 7,
	LZ4B200_ST_OFFSET_TRUNCATED    = 8,  /* block ends inside the 2-byte offset (ref: Constraint_Error) */
	LZ4B200_ST_OUTPUT_OVERFLOW     = 9,  /* would write past dst_cap (ref: unchecked write) */
	LZ4B200_ST_NEEDS_HISTORY       = 10, /* not an error: a block decoded independently reaches into
	                                      * the previous block; the host re-runs the frame through
	                                      * lz4b200_decode_linked (reference accepts such frames) */
	LZ4B200_ST_NOT_RUN             = 11  /* linked frame: an earlier block of the frame failed */
};

/* A run of consecutive block descriptors decoded strictly in order by one warp, each block's
 * output starting where the previous one ended: one linked frame, or a whole stream of frames
 * that needs exact placement (LZ4B200_BLK_FIRST_OF_FRAME restarts the history). */
typedef struct lz4b200_chain {
	uint32_t first_block;
	uint32_t n_blocks;
	uint64_t dst_off;   /* output of the chain starts here ... */
	uint64_t dst_cap;   /* ... and may not grow beyond this many bytes */
} lz4b200_chain;

/* The blocks of one frame, for the content checksum over their concatenated output. */
typedef struct lz4b200_frame_blocks {
	uint32_t first_block;
	uint32_t n_blocks;
} lz4b200_frame_blocks;

/* A byte range to hash (content checksum of one frame), relative to `data`. */
typedef struct lz4b200_hash_span {
	uint64_t off;
	uint64_t len;
} lz4b200_hash_span;

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t the
 * caller owns (e.g. torch's current stream) or NULL to let the context create
 * its own non-blocking stream.  A context is used by one host thread at a time. */
int lz4b200_create(int device, void *stream, lz4b200_ctx **out);
/* A context is reference counted.  lz4b200_create returns it with one reference (the caller's);
 * lz4b200_destroy drops one, and the CUDA objects go when the last one is dropped.  Every stream
 * object, decompressor and batch made on a context holds its own reference, so destroying the
 * context before its children is safe: they keep working and the teardown happens with the last
 * of them (the shape an Ada Limited_Controlled finaliser needs -- finalisation order of a
 * Decompressor and its device context is not the caller's to choose).  lz4b200_retain takes an
 * additional reference for a holder of its own. */
int lz4b200_retain(lz4b200_ctx *ctx);
int lz4b200_destroy(lz4b200_ctx *ctx);
/* Text of the last CUDA failure seen by this context (never NULL). */
const char *lz4b200_last_error(const lz4b200_ctx *ctx);
/* SM count of the context's device (grid sizing is a multiple of it). */
int lz4b200_sm_count(const lz4b200_ctx *ctx);
/* Number of kernel launches issued by this context since creation. */
uint64_t lz4b200_launch_count(const lz4b200_ctx *ctx);

/* K1 tuning: which generation of the independent-block kernel lz4b200_decode_blocks launches.
 *    0        choose from the block count (default): v6 from 20 000 blocks on (a launch of v6 lasts as long as one
 *             lane's block whatever the count, v4 costs ~0.37 us per 64 KiB text block), v4 below; the batch
 *             scheduler, which has the block table, counts only the blocks that take long
 *   60 / 61   v6: one lane per block, one piece of a sequence (<= 8 literal + <= 16 match bytes) per trip, the parse
 *             K pieces ahead of the copy; 60 = 256-byte out ring per lane, 14 warps per SM; 61 = 512-byte ring, 8 warps
 *   50        v5: one lane per block, a whole sequence per trip (round 1's kernel, kept for A/B)
 *   40        v4: one warp per block, warp-parallel parse, 4 KiB shared-memory output ring per warp;
 *             41 / 42 / 44 / 48 fix the number of blocks a warp hashes together and decodes in turn
 *   64        v3: one CTA per block, whole output window in shared memory (kept for A/B: slow)
 *   1..16     v2: that many blocks side by side per warp, output assembled in global memory
 *   -1        v1: one warp per block, every lane in lock-step; also the exact fallback of all others
 * All of them produce identical bytes and statuses (tests/test_gpu_parity.py runs every one). */
int lz4b200_set_tuning(lz4b200_ctx *ctx, int blocks_per_warp);

int lz4b200_get_tuning(const lz4b200_ctx *ctx);

/* Statistics of the last lane-per-block K1 launch (v6) on the context's current lane: how many blocks its fast path
 * handed to the exact routine, and how many of those the idle-trip safety net sent there.  Synchronises the lane. */
int lz4b200_k1_fallbacks(lz4b200_ctx *ctx, uint32_t *to_exact, uint32_t *by_safety_net);

/* Name of the kernel lz4b200_decode_blocks would launch for n_blocks under the current tuning
 * (for logs and benchmark records). */
const char *lz4b200_k1_kernel_name(const lz4b200_ctx *ctx, uint32_t n_blocks);

/* A context owns up to four CUDA streams ("lanes"; lane 0 is the primary one given to / made by
 * lz4b200_create).  Every call enqueues on the lane selected last.  The batch scheduler uses them
 * to overlap H2D of one chunk, kernels of another and D2H of a third; no kernel waits on another
 * lane.  lz4b200_sync_all waits for all lanes. */
int lz4b200_use_lane(lz4b200_ctx *ctx, int lane);
int lz4b200_sync_all(lz4b200_ctx *ctx);

/* Device and pinned-host memory, so that the Ada side never links libcudart. */
int lz4b200_alloc(lz4b200_ctx *ctx, size_t bytes, void **dev_ptr);
int lz4b200_free(lz4b200_ctx *ctx, void *dev_ptr);
int lz4b200_alloc_host(lz4b200_ctx *ctx, size_t bytes, void **host_ptr);
int lz4b200_free_host(lz4b200_ctx *ctx, void *host_ptr);
/* Asynchronous on the context stream; pair with lz4b200_sync. */
int lz4b200_h2d(lz4b200_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int lz4b200_d2h(lz4b200_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int lz4b200_memset(lz4b200_ctx *ctx, void *dst_dev, int value, size_t bytes);
int lz4b200_sync(lz4b200_ctx *ctx);

/* Device-side timing on the context stream (CUDA events). */
int lz4b200_timer_start(lz4b200_ctx *ctx);
int lz4b200_timer_stop(lz4b200_ctx *ctx, float *elapsed_ms);   /* synchronises */

/* The roofline denominator measured in place: a plain grid-stride copy of `bytes` from src_dev to dst_dev (read +
 * write = 2 * bytes of HBM traffic), `reps` timed passes after one warm-up; best_ms = the fastest pass. */
int lz4b200_copy_probe(lz4b200_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes, int reps, float *best_ms);
/* CUDA device ordinal of the context. */
int lz4b200_device_of(const lz4b200_ctx *ctx);

/* CUDA events on the context stream, for per-kernel timing without libcudart on the caller's
 * side.  elapsed is valid once the stream has been synchronised past `stop`. */
int lz4b200_event_create(lz4b200_ctx *ctx, void **event);
int lz4b200_event_destroy(lz4b200_ctx *ctx, void *event);
int lz4b200_event_record(lz4b200_ctx *ctx, void *event);
int lz4b200_event_sync(lz4b200_ctx *ctx, void *event);   /* the host waits until the lane has passed the record */
int lz4b200_event_elapsed(lz4b200_ctx *ctx, void *start, void *stop, float *elapsed_ms);

/* K1 (+K2): decode `n_blocks` mutually independent blocks, one warp per block,
 * block XXH32 fused.  Takes over Decode_Full_Block_With_Trailer,
 * Check_Checksum, Decompress_Full_Block, Write_Output and Output_With_History
 * (lib/lz4ada.adb:661-904).  All pointers are device pointers. */
int lz4b200_decode_blocks(lz4b200_ctx *ctx, const uint8_t *src, uint8_t *dst,
		uint32_t n_blocks, const lz4b200_blk_desc *desc,
		lz4b200_blk_status *status);

/* K2: stored (uncompressed) blocks as a wide copy (lib/lz4ada.adb:685-695): block idx[k] of the table for k <
 * n_idx, every block cut into 32 KiB tiles with a warp each (max_len = the longest of them).  spans / scratch (n_idx
 * entries each, or NULL when no block carries a checksum): spans[k] = the payload of block idx[k] in src; its XXH32
 * is computed as a chain of its own beside the copy (Check_Checksum, :698-707) and compared with the trailer. */
int lz4b200_copy_stored(lz4b200_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t n_idx, const uint32_t *idx,
		uint32_t max_len, const lz4b200_blk_desc *desc, lz4b200_blk_status *status,
		const lz4b200_hash_span *spans, uint32_t *scratch);

/* Chains -- the blocks of a chain in order, chains in parallel; the running output
 * position of the chain places every block, so matches may reach back across block
 * boundaries of the same frame (linked frames; big independent blocks as chains of one;
 * also the exact-placement retry path).  Replaces the ring/history handling of
 * lib/lz4ada.adb:678-680, 845-904.  Kernel: K7 (one CTA per chain: speculative parallel
 * parse + pointer jumping in a shared-memory window, kernels_k7.cuh), with the block
 * checksums of all chained blocks hashed by a launch of their own in front of it.
 * LZ4B200_CHAIN_KERNEL=pipe | k6 | warp in the environment selects the earlier generations
 * (K4 pipeline, K6 rounds, one warp per chain) for A/B; LZ4B200_K7_DEBUG=1 prints K7's
 * phase statistics after every launch. */
int lz4b200_decode_linked(lz4b200_ctx *ctx, const uint8_t *src, uint8_t *dst,
		uint32_t n_chains, const lz4b200_chain *chains,
		const lz4b200_blk_desc *desc, lz4b200_blk_status *status);

/* Statistics of the chain kernel K7 since the last call (per device, all contexts): blocks its fast path finished, and
 * blocks it gave up on (that block and the rest of its chain went to the exact routine).  Synchronises the lane; resets. */
int lz4b200_chain_stats(lz4b200_ctx *ctx, uint32_t *finished, uint32_t *given_up);

/* K3: XXH32 (seed 0) of `n` byte ranges, one serial chain per range, ranges in
 * parallel.  Takes over Update_Checksum / XXHash32.Update / Final for the
 * content checksum (lib/lz4ada.adb:709-714, 942-1017). */
int lz4b200_xxh32_spans(lz4b200_ctx *ctx, const uint8_t *data, uint32_t n,
		const lz4b200_hash_span *spans, uint32_t *out);

/* K3 for the batch path: content checksum of each frame straight after K1/K4 without a
 * host round trip -- the length of a frame's output is summed from the block statuses on
 * the device.  valid[f] = 0 when a block of the frame failed or the blocks' outputs are not
 * contiguous (the host then re-places the frame and asks again). */
int lz4b200_xxh32_frames(lz4b200_ctx *ctx, const uint8_t *dst, uint32_t n_frames,
		const lz4b200_frame_blocks *frames, const lz4b200_blk_desc *desc,
		const lz4b200_blk_status *status, uint32_t *digest, uint32_t *valid);

/* K5: size pre-pass -- walks the sequences of each block without writing
 * output and reports out_len (and structural errors) in status.  Used when a
 * frame's interior blocks are not all block-max sized (the frame format
 * carries no per-block decompressed size). */
int lz4b200_size_blocks(lz4b200_ctx *ctx, const uint8_t *src, uint32_t n_blocks,
		const lz4b200_blk_desc *desc, lz4b200_blk_status *status);

/* Synchronous single-block path under Decompressor.Update.  A stream object owns a
 * device-resident window [64 KiB history | block output ...] and the streaming XXH32 state
 * of the content checksum, so nothing but the new block crosses PCIe.  There is no CPU decode
 * behind it.  max_block = largest payload / output of one call (block max + 64 KiB + 8 for a
 * caller honouring Min_Buffer_Size). */
typedef struct lz4b200_stream lz4b200_stream;
int lz4b200_stream_create(lz4b200_ctx *ctx, uint32_t max_block, lz4b200_stream **out);
int lz4b200_stream_destroy(lz4b200_stream *s);
/* New frame: forget history, reset the content hash (Reset_Outer_For_Next_Frame, :451-461). */
int lz4b200_stream_reset(lz4b200_stream *s);
/* H2D of one block (payload + optional 4-byte checksum trailer), one warp decodes it behind
 * the window's history, optional streaming content-hash update over the produced bytes,
 * D2H of the produced bytes to host_dst.  flags = LZ4B200_BLK_*. */
int lz4b200_stream_block(lz4b200_stream *s, const uint8_t *host_src, uint32_t src_len,
		uint32_t flags, int hash_content, uint8_t *host_dst, uint32_t dst_cap,
		lz4b200_blk_status *status);
/* The same with a hint: expect_out = what a well-formed block of this frame produces at most (the frame's block
 * maximum).  Up to 256 KiB the call then needs ONE synchronisation instead of two: the content hash reads the block's
 * status on the device and expect_out bytes come back together with the status (bytes of host_dst behind the block's
 * output and below expect_out are overwritten -- free space of the caller's Buffer).  0 = no hint. */
int lz4b200_stream_block2(lz4b200_stream *s, const uint8_t *host_src, uint32_t src_len,
		uint32_t flags, int hash_content, uint8_t *host_dst, uint32_t dst_cap,
		uint32_t expect_out, lz4b200_blk_status *status);
/* XXH32 of every byte produced since the last reset (XXHash32.Final, :993-1017). */
int lz4b200_stream_digest(lz4b200_stream *s, uint32_t *xxh32);
/* Read-ahead support: n decoded bytes that already sit in device memory
 * (a block decoded ahead of time by lz4b200_decode_blocks into a staging buffer) become the next
 * bytes of the stream -- appended to its history window and, if hash_content, folded into its
 * running content checksum, exactly what lz4b200_stream_block does for a block it decodes itself.
 * Asynchronous on the context's current stream. */
int lz4b200_stream_adopt(lz4b200_stream *s, const uint8_t *dev_bytes, uint32_t n, int hash_content);

/* ... the same for up to 255 blocks at once, in stream order: piece i = lengths[i] bytes at dev_base + offsets[i].
 * One hash launch for all of them (the running content checksum is one serial chain anyway) and the last 64 KiB as
 * the new history window, instead of a copy and a launch per block.  Asynchronous on the stream's lane; the pieces
 * must stay intact until the lane has passed them. */
int lz4b200_stream_adopt_list(lz4b200_stream *s, const uint8_t *dev_base, uint32_t n_pieces, const uint32_t *offsets,
		const uint32_t *lengths, int hash_content);

/* ------------------------------------------------------------------------
 * LZ4Ada package API  (reference lib/lz4ada.ads)
 * --------------------------------------------------------------------- */

/* Flexible_Memory_Reservation, lib/lz4ada.ads:79-80 (same order) */
enum lz4ada_reservation {
	LZ4ADA_SZ_64_KIB = 0, LZ4ADA_SZ_256_KIB, LZ4ADA_SZ_1_MIB, LZ4ADA_SZ_4_MIB,
	LZ4ADA_SZ_8_MIB, LZ4ADA_USE_FIRST, LZ4ADA_SINGLE_FRAME
};
#define LZ4ADA_FOR_MODERN LZ4ADA_SZ_4_MIB   /* lib/lz4ada.ads:92  */
#define LZ4ADA_FOR_LEGACY LZ4ADA_SZ_8_MIB   /* lib/lz4ada.ads:100 */
#define LZ4ADA_FOR_ALL    LZ4ADA_SZ_8_MIB   /* lib/lz4ada.ads:106 */

/* End_Of_Frame, lib/lz4ada.ads:124 */
enum lz4ada_end_of_frame { LZ4ADA_EOF_YES = 0, LZ4ADA_EOF_NO = 1, LZ4ADA_EOF_MAYBE = 2 };

/* The five exceptions of lib/lz4ada.ads:133-162, plus the non-library outcomes. */
enum lz4ada_exception {
	LZ4ADA_OK = 0,
	LZ4ADA_CHECKSUM_ERROR,
	LZ4ADA_DATA_CORRUPTION,
	LZ4ADA_NOT_SUPPORTED,
	LZ4ADA_TOO_FEW_HEADER_BYTES,
	LZ4ADA_TOO_LITTLE_MEMORY,
	LZ4ADA_CONSTRAINT_ERROR,   /* "Library bug detected", lib/lz4ada.adb:185 */
	LZ4ADA_ASSERTION_ERROR,    /* violated precondition (Pre => ...) */
	LZ4ADA_DEVICE_ERROR        /* CUDA failure: distinct from every LZ4 error */
};

typedef struct lz4ada_decompressor lz4ada_decompressor;

/* Bind the LZ4Ada layer of this thread/process to a device context.  Every
 * decompressor created afterwards decodes its blocks on that device.  Passing
 * NULL makes the library create (once) a context on device 0. */
int lz4ada_set_device_context(lz4b200_ctx *ctx);

/* Init, lib/lz4ada.ads:189/218, lib/lz4ada.adb:48-63 */
int lz4ada_init(int *min_buffer_size, int reservation, lz4ada_decompressor **out);
/* Init_With_Header, lib/lz4ada.ads:238, lib/lz4ada.adb:79-125.  Pre: input_len >= 7. */
int lz4ada_init_with_header(const uint8_t *input, int input_len, int *num_consumed,
		int *min_buffer_size, int reservation, lz4ada_decompressor **out,
		char *message, size_t message_cap);
/* Init_For_Block, lib/lz4ada.ads:255, lib/lz4ada.adb:127-147 */
int lz4ada_init_for_block(int *min_buffer_size, int compressed_length, int reservation,
		lz4ada_decompressor **out);
/* Update (Octets flavour, Buffer'First = 0), lib/lz4ada.ads:281, lib/lz4ada.adb:383-418.
 * One step per call exactly as the reference. */
int lz4ada_update(lz4ada_decompressor *ctx, const uint8_t *input, int input_len,
		int *num_consumed, uint8_t *buffer, int buffer_len,
		int *output_first, int *output_last);
/* Is_End_Of_Frame, lib/lz4ada.ads:303 */
int lz4ada_is_end_of_frame(const lz4ada_decompressor *ctx);
/* "raised LZ4ADA.<NAME> : <message>" of the last failing call on ctx. */
const char *lz4ada_exception_message(const lz4ada_decompressor *ctx);
void lz4ada_free(lz4ada_decompressor *ctx);

/* To_Hex, lib/lz4ada.ads:306-307 (lower case, zero padded; out needs 3 / 9 bytes) */
void lz4ada_to_hex_u8(uint8_t num, char *out);
void lz4ada_to_hex_u32(uint32_t num, char *out);

/* package XXHash32, lib/lz4ada.ads:311-321.  Host-side like the reference's:
 * it serves the frame-header checksum and callers such as tool_xxhash32ada;
 * block and content checksums of the decode path are computed on the device. */
typedef struct lz4ada_xxhash32 {
	uint32_t state[4];
	uint8_t  buffer[16];
	int32_t  buffer_size;
	uint64_t total_length;
} lz4ada_xxhash32;
void     lz4ada_xxhash32_init(lz4ada_xxhash32 *h, uint32_t seed);   /* NB: ignores seed like :925-930 */
void     lz4ada_xxhash32_reset(lz4ada_xxhash32 *h, uint32_t seed);
void     lz4ada_xxhash32_update(lz4ada_xxhash32 *h, const uint8_t *input, size_t len);
uint32_t lz4ada_xxhash32_final(const lz4ada_xxhash32 *h);
uint32_t lz4ada_xxhash32_hash(const uint8_t *input, size_t len);

/* ------------------------------------------------------------------------
 * Batched device entry point (added by this library; north star)
 * --------------------------------------------------------------------- */

/* One input stream = what a caller would feed to Init(For_All) + Update until
 * end of input: one or more concatenated modern / legacy / skippable frames. */
typedef struct lz4ada_batch_item {
	uint64_t src_off;   /* stream start within the batch source buffer */
	uint64_t src_len;
	uint64_t dst_off;   /* where this stream's output goes in the batch output buffer */
	uint64_t dst_cap;   /* 0 = let the planner place it (packed, 256-byte aligned) */
} lz4ada_batch_item;

typedef struct lz4ada_batch_result {
	int32_t  exception;     /* enum lz4ada_exception of the first error in stream order */
	int32_t  end_of_frame;  /* Is_End_Of_Frame after the last byte */
	uint32_t n_frames;      /* frames seen (including skippable) */
	uint32_t n_blocks;
	uint64_t dst_off;       /* output placement actually used */
	uint64_t out_len;       /* bytes produced before any error */
} lz4ada_batch_result;

typedef struct lz4ada_batch lz4ada_batch;

#define LZ4ADA_BATCH_SRC_ON_DEVICE 1u  /* src_dev already holds the compressed bytes */

/* Host stage: parse every frame header and walk the block size words
 * (lib/lz4ada.adb:155-361, 525-585) to build the block table.  Needs the
 * compressed bytes in host memory (`src_host`); nothing is decoded here.
 * ctx may be NULL (the process-wide default context is taken at upload).
 * reservation: LZ4ADA_SZ_64_KIB .. LZ4ADA_SZ_8_MIB decode every stream as Init(reservation) + Update
 * would; LZ4ADA_USE_FIRST / LZ4ADA_SINGLE_FRAME as Init_With_Header(stream, reservation) + Update
 * (lib/lz4ada.adb:79-125; the call tool_unlz4ada and Test_Error_Case make), including
 * Too_Few_Header_Bytes and the Single_Frame policing. */
int lz4ada_batch_plan(lz4b200_ctx *ctx, const uint8_t *src_host, uint64_t src_bytes,
		uint32_t n_items, const lz4ada_batch_item *items, int reservation,
		lz4ada_batch **out);
/* Inspection of the plan (tests, tooling): block `index` of the table in stream order, and
 * what the host stage alone concluded about stream `item` (header / size-word / Single_Frame
 * errors, Is_End_Of_Frame at end of input).  Pure host work: usable without a device. */
int lz4ada_batch_block_desc(const lz4ada_batch *b, uint64_t index, lz4b200_blk_desc *out);
int lz4ada_batch_host_outcome(const lz4ada_batch *b, uint32_t item, int *exception, int *end_of_frame,
		uint32_t *n_frames, uint32_t *n_blocks, char *message, size_t message_cap);
/* Output bytes the plan needs (for allocating the destination). */
uint64_t lz4ada_batch_output_bytes(const lz4ada_batch *b);
uint64_t lz4ada_batch_block_count(const lz4ada_batch *b);
/* Algorithmic traffic of one run: compressed bytes read + bytes written
 * + bytes re-read for content checksums. */
void lz4ada_batch_traffic(const lz4ada_batch *b, uint64_t *compressed_read,
		uint64_t *decompressed_written, uint64_t *checksum_reread);
/* Device time of the kernels of the last lz4ada_batch_run, from CUDA events recorded on the
 * launching stream: ms[0] = K1 (independent blocks), ms[1] = K4 (chains), ms[2] = K3 (content
 * checksums).  0 for a kernel that did not run. */
int lz4ada_batch_kernel_ms(const lz4ada_batch *b, float ms[3]);

/* How many bytes the caller's output buffer (dst_dev) really holds, when that is more than
 * lz4ada_batch_output_bytes.  The plan gives every block one block maximum of room, which is what every
 * well-formed frame needs.  The reference, however, bounds a block's output only by its caller's Buffer
 * (Min_Buffer_Size = reservation + 64 KiB + 8, lib/lz4ada.adb:54, 813-820), so a crafted frame whose block
 * inflates past the declared block maximum decodes there (test vector cntblkszoverflow under Init(For_All)).
 * Such a stream is decoded again as a chain under the reference's bound; when it no longer fits its region
 * it is placed behind the planned output, in [output_bytes, capacity), and its result reports that dst_off.
 * Without spare capacity it ends with Data_Corruption "Output buffer exhausted ..." naming its region. */
int lz4ada_batch_set_output_capacity(lz4ada_batch *b, uint64_t bytes);

/* Exact sizing: call between lz4ada_batch_plan and lz4ada_batch_upload.
 * The upload then runs the size pre-pass K5 over every block of the independent-block frames and
 * places the blocks back to back at their true sizes instead of one block-maximum apart: the output
 * of a stream is contiguous from the first run on, so frames with short interior blocks (writers that
 * flush often) never need the decode-again-as-a-chain retry.  lz4ada_batch_output_bytes keeps
 * reporting the upper bound the caller allocated from.  Costs one extra pass over the compressed
 * bytes (one thread per block). */
int lz4ada_batch_exact_sizing(lz4ada_batch *b);
/* Name of the K1 kernel the last lz4ada_batch_run launched (the batch scheduler overrides K1's
 * block-count rule when too few of the blocks are long-running ones). */
const char *lz4ada_batch_k1_kernel_name(const lz4ada_batch *b);
/* How many streams the last run had to decode a second time as chains with exact placement. */
uint32_t lz4ada_batch_retried_streams(const lz4ada_batch *b);
/* Device stage: upload the tables (and the compressed bytes unless they are
 * already on the device), run K1..K4, fetch statuses and fold them in stream
 * order.  src_dev / dst_dev are device pointers sized src_bytes(+32 slack) /
 * lz4ada_batch_output_bytes(+32 slack).  May be called repeatedly. */
int lz4ada_batch_upload(lz4ada_batch *b, const uint8_t *src_host, uint8_t *src_dev);
int lz4ada_batch_run(lz4ada_batch *b, const uint8_t *src_dev, uint8_t *dst_dev);
int lz4ada_batch_results(const lz4ada_batch *b, lz4ada_batch_result *results);
const char *lz4ada_batch_message(const lz4ada_batch *b, uint32_t item);
void lz4ada_batch_free(lz4ada_batch *b);

/* Host buffers in and out with the device stage cut into `n_chunks` groups of streams that are
 * pipelined over the context's lanes: H2D of chunk k+1 and D2H of chunk k-1 overlap the kernels of
 * chunk k.  src_dev / dst_dev are device scratch buffers sized like lz4ada_batch_upload / _run. */
int lz4ada_batch_run_pipelined(lz4ada_batch *b, const uint8_t *src_host, uint8_t *dst_host,
		uint8_t *src_dev, uint8_t *dst_dev, uint32_t n_chunks);

/* One call, host buffers in and out: plan + H2D + kernels + D2H.  This is the
 * end-to-end call bench.py times for `e2e`.  items[k].dst_off / dst_cap are updated with
 * the placement used; `messages` (optional) receives n_items strings of message_stride bytes. */
int lz4ada_batch_decompress(lz4b200_ctx *ctx, const uint8_t *src_host, uint64_t src_bytes,
		uint8_t *dst_host, uint64_t dst_bytes, uint32_t n_items,
		lz4ada_batch_item *items, int reservation, lz4ada_batch_result *results,
		char *messages, size_t message_stride);

/* Name of the K1 kernel the last lz4ada_batch_decompress on this context launched for its chunks. */
const char *lz4ada_last_k1_kernel_name(lz4b200_ctx *ctx);

/* The same call over several GPUs of one box: whole streams are dealt to the contexts (one per
 * GPU) in contiguous runs balanced by compressed size -- a content-checksummed frame is never split, nothing is exchanged between devices
 * -- and every context runs its share on a host thread of its own (plan, H2D, kernels, D2H), all reading from and
 * writing into the caller's two host buffers.  Results, messages and the placement written back to items[] are in
 * the order the streams were given.  n_ctx = 1 is lz4ada_batch_decompress. */
int lz4ada_batch_decompress_multi(uint32_t n_ctx, lz4b200_ctx *const *ctxs, const uint8_t *src_host, uint64_t src_bytes,
		uint8_t *dst_host, uint64_t dst_bytes, uint32_t n_items, lz4ada_batch_item *items, int reservation,
		lz4ada_batch_result *results, char *messages, size_t message_stride);

#ifdef __cplusplus
}
#endif
#endif /* LZ4B200_H */
