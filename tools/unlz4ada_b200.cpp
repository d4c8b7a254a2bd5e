// unlz4ada_b200 -- stdin -> stdout decompressor over liblz4b200.so, the counterpart of the
// reference's tool_unlz4ada_simple/unlz4ada_simple.adb:23-36 (Init + Update loop on 4 KiB reads;
// the library handles concatenated modern / legacy / skippable frames).  Every block is decoded on
// the GPU; without a device the tool fails with the library's DEVICE_ERROR text.
//   exit 0: ok      exit 1: LZ4Ada exception (text on stderr, like GNAT's unhandled-exception line)
//   exit 2: input ended mid-frame (the reference raises Constraint_Error "Input ended mid-frame.")
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lz4b200.h"

int main()
{
	int min_buffer_size = 0;
	lz4ada_decompressor *ctx = nullptr;
	if (lz4ada_init(&min_buffer_size, LZ4ADA_FOR_ALL, &ctx) != LZ4ADA_OK) return 3;
	std::vector<uint8_t> output(static_cast<size_t>(min_buffer_size));
	uint8_t input[4096];
	for (;;) {
		const size_t got = fread(input, 1, sizeof input, stdin);
		if (got == 0) break;
		size_t pos = 0;
		while (pos < got) {
			int consumed = 0, first = 1, last = 0;
			const int rc = lz4ada_update(ctx, input + pos, static_cast<int>(got - pos), &consumed, output.data(),
						     min_buffer_size, &first, &last);
			if (rc != LZ4ADA_OK) {
				fprintf(stderr, "%s\n", lz4ada_exception_message(ctx));
				return 1;
			}
			if (last >= first) fwrite(output.data() + first, 1, static_cast<size_t>(last - first + 1), stdout);
			pos += static_cast<size_t>(consumed);
		}
	}
	const int eof = lz4ada_is_end_of_frame(ctx);
	lz4ada_free(ctx);
	fflush(stdout);
	if (eof == LZ4ADA_EOF_NO) {
		fprintf(stderr, "raised CONSTRAINT_ERROR : Input ended mid-frame.\n");
		return 2;
	}
	return 0;
}
