// unlz4ada_b200 -- stdin -> stdout decompressor over liblz4b200.so, the counterpart of the reference's
// tool_unlz4ada/unlz4ada.adb:64-105 (and tool_unlz4ada_simple): concatenated modern / legacy / skippable frames in,
// plain bytes out.  Every block is decoded on the GPU; without a device the tool fails with the library's
// DEVICE_ERROR text.
//
//   unlz4ada_b200 [-v]            (default) the whole input goes to the batched device entry point in one call
//                                 (lz4ada_batch_decompress: block table on the host, H2D, kernels, D2H) -- SURVEY.md 8f-1
//   unlz4ada_b200 --update [-v]   the drop-in streaming API instead: Init + Update on 32 MiB reads, one block per call
//                                 (tool_unlz4ada_simple/unlz4ada_simple.adb:23-36 with a bigger read)
//   --file F --repeat N           read F instead of stdin, N passes in one process (the first pays for the CUDA context)
//   --keep                        with --update --file --repeat: ONE decompressor for all passes (the file's frames N
//                                 times over as concatenated frames): passes 2.. show the steady state of a long
//                                 stream, without the device buffers a new decompressor allocates
//   -v                            timing on stderr, I/O included: bytes in / out, seconds from the first read to the last
//                                 write, decompressed MB/s; and the same without I/O
//   exit 0: ok      exit 1: LZ4Ada exception (text on stderr, like GNAT's unhandled-exception line; the bytes decoded
//   before it are written first, as the reference's tools do)      exit 2: input ended mid-frame
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lz4b200.h"

static double now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static lz4ada_decompressor *g_kept = nullptr;   // --keep
static int g_kept_min = 0;

static int run_update(bool verbose, FILE *in_f, FILE *out_f, bool keep = false)
{
	const double t0 = now();
	int min_buffer_size = 0;
	lz4ada_decompressor *ctx = nullptr;
	if (keep && g_kept) {
		ctx = g_kept;
		min_buffer_size = g_kept_min;
	} else if (lz4ada_init(&min_buffer_size, LZ4ADA_FOR_ALL, &ctx) != LZ4ADA_OK) {
		return 3;
	}
	std::vector<uint8_t> output(static_cast<size_t>(min_buffer_size));
	std::vector<uint8_t> input(32u << 20);   // big reads: the read-ahead decodes what is complete in one Input
	size_t total_in = 0, total_out = 0;
	double t_lib = 0;
	for (;;) {
		// big reads: the library decodes ahead when it sees whole blocks behind the current one in its Input
		const size_t got = fread(input.data(), 1, input.size(), in_f);
		if (got == 0) break;
		total_in += got;
		size_t pos = 0;
		while (pos < got) {
			int consumed = 0, first = 1, last = 0;
			const double a = now();
			const int rc = lz4ada_update(ctx, input.data() + pos, static_cast<int>(got - pos), &consumed, output.data(),
						     min_buffer_size, &first, &last);
			t_lib += now() - a;
			if (rc != LZ4ADA_OK) {
				if (out_f) fflush(out_f);
				fprintf(stderr, "%s\n", lz4ada_exception_message(ctx));
				return 1;
			}
			if (last >= first) {
				if (out_f) fwrite(output.data() + first, 1, static_cast<size_t>(last - first + 1), out_f);
				total_out += static_cast<size_t>(last - first + 1);
			}
			pos += static_cast<size_t>(consumed);
		}
	}
	const int eof = lz4ada_is_end_of_frame(ctx);
	if (keep) {
		g_kept = ctx;
		g_kept_min = min_buffer_size;
	} else {
		lz4ada_free(ctx);
	}
	if (out_f) fflush(out_f);
	if (verbose) {
		const double dt = now() - t0;
		fprintf(stderr, "update: %zu bytes in, %zu bytes out, %.3f s with I/O = %.1f MB/s decompressed; %.3f s inside Update = %.1f MB/s\n",
			total_in, total_out, dt, total_out / dt / 1e6, t_lib, total_out / t_lib / 1e6);
	}
	if (eof == LZ4ADA_EOF_NO) {
		fprintf(stderr, "raised CONSTRAINT_ERROR : Input ended mid-frame.\n");
		return 2;
	}
	return 0;
}

static int run_batch(bool verbose, FILE *in_f, FILE *out_f)
{
	const double t0 = now();
	std::vector<uint8_t> in;
	{
		std::vector<uint8_t> chunk(8u << 20);
		for (;;) {
			const size_t got = fread(chunk.data(), 1, chunk.size(), in_f);
			if (got == 0) break;
			in.insert(in.end(), chunk.begin(), chunk.begin() + got);
		}
	}
	if (in.empty()) return 0;
	in.resize(in.size() + 64);   // slack the device copy may read
	const size_t n_in = in.size() - 64;
	const double t1 = now();
	// plan on the host to learn the output room (nothing is decoded here), then one call for everything
	lz4ada_batch_item item;
	memset(&item, 0, sizeof item);
	item.src_len = n_in;
	lz4ada_batch *plan = nullptr;
	if (lz4ada_batch_plan(nullptr, in.data(), n_in, 1, &item, LZ4ADA_FOR_ALL, &plan) != LZ4ADA_OK) return 3;
	const uint64_t need = lz4ada_batch_output_bytes(plan);
	lz4ada_batch_free(plan);
	const uint64_t cap = need + (16u << 20);
	std::vector<uint8_t> out(cap + 64);
	lz4ada_batch_result res;
	memset(&res, 0, sizeof res);
	char msg[512] = "";
	const int rc = lz4ada_batch_decompress(nullptr, in.data(), n_in, out.data(), cap, 1, &item, LZ4ADA_FOR_ALL, &res, msg, sizeof msg);
	const double t2 = now();
	if (rc != LZ4ADA_OK) {
		fprintf(stderr, "raised LZ4ADA.DEVICE_ERROR : lz4ada_batch_decompress failed (rc=%d): no CUDA device or driver. This library has no CPU decode path.\n", rc);
		return 1;
	}
	if (res.out_len && out_f) fwrite(out.data() + res.dst_off, 1, res.out_len, out_f);
	if (out_f) fflush(out_f);
	if (verbose) {
		const double dt = now() - t0;
		fprintf(stderr, "batch: %zu bytes in, %llu bytes out, %.3f s with I/O = %.1f MB/s decompressed; %.3f s in the batch call (table + H2D + kernels + D2H) = %.1f MB/s\n",
			n_in, static_cast<unsigned long long>(res.out_len), dt, res.out_len / dt / 1e6, t2 - t1, res.out_len / (t2 - t1) / 1e6);
	}
	if (res.exception != LZ4ADA_OK) {
		fprintf(stderr, "%s\n", msg);
		return 1;
	}
	if (res.end_of_frame == LZ4ADA_EOF_NO) {
		fprintf(stderr, "raised CONSTRAINT_ERROR : Input ended mid-frame.\n");
		return 2;
	}
	return 0;
}

int main(int argc, char **argv)
{
	bool update = false, verbose = false, keep = false;
	const char *file = nullptr;
	int repeat = 1;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--update")) update = true;
		else if (!strcmp(argv[i], "-v")) verbose = true;
		else if (!strcmp(argv[i], "--keep")) keep = true;
		else if (!strcmp(argv[i], "--file") && i + 1 < argc) file = argv[++i];
		else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
		else {
			fprintf(stderr, "usage: unlz4ada_b200 [--update] [-v] [--file in.lz4 [--repeat N]] < in.lz4 > out\n");
			return 3;
		}
	}
	if (!file) return update ? run_update(verbose, stdin, stdout) : run_batch(verbose, stdin, stdout);
	// --file F --repeat N: decode F N times in this process (the first pass pays for the CUDA context, the device
	// and pinned buffers; later passes show the path itself); only the last pass writes its output
	int rc = 0;
	for (int r = 0; r < repeat && rc == 0; r++) {
		FILE *f = fopen(file, "rb");
		if (!f) { perror(file); return 3; }
		FILE *out_f = r + 1 == repeat ? stdout : nullptr;
		rc = update ? run_update(verbose, f, out_f, keep) : run_batch(verbose, f, out_f);
		fclose(f);
	}
	if (g_kept) lz4ada_free(g_kept);
	return rc;
}
