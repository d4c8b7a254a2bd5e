/*
 * lz4_greedy_enc.c -- small greedy LZ4 *block encoder* for the test / bench harness.
 * Used only to make synthetic inputs when liblz4.so.1 cannot be loaded (BASELINE.json
 * north_star: "otherwise with a small harness-side greedy LZ4 encoder").  Not on the decode path.
 *
 * lz4g_compress(buf, prefix_len, n, dst, cap): encodes buf[prefix_len .. prefix_len+n) as one LZ4
 * block; matches may reach back into the prefix (linked blocks), at most 65535 bytes.
 * Follows the block-format end rules: the last 5 bytes are literals and the last match starts
 * at least 12 bytes before the end of the block.
 */
#include <stdint.h>
#include <string.h>

#define HASH_BITS 16
#define MINMATCH 4
#define MFLIMIT 12
#define LASTLITERALS 5

int lz4g_bound(int n) { return n + n / 255 + 16; }

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint32_t hash4(uint32_t v) { return (v * 2654435761u) >> (32 - HASH_BITS); }

static uint8_t *put_len(uint8_t *op, int len)
{
	while (len >= 255) { *op++ = 255; len -= 255; }
	*op++ = (uint8_t)len;
	return op;
}

int lz4g_compress(const uint8_t *buf, int prefix_len, int n, uint8_t *dst, int cap)
{
	static __thread int32_t table[1 << HASH_BITS];
	const uint8_t *base = buf;
	const uint8_t *start = buf + prefix_len;
	const uint8_t *end = start + n;
	const uint8_t *ip = start, *anchor = start;
	const uint8_t *mflimit = end - MFLIMIT, *matchlimit = end - LASTLITERALS;
	uint8_t *op = dst, *oend = dst + cap;
	int i;
	for (i = 0; i < (1 << HASH_BITS); i++) table[i] = -1;
	for (i = 0; i + MINMATCH <= prefix_len; i++) table[hash4(rd32(base + i))] = i;
	if (n >= MFLIMIT + 1) {
		while (ip <= mflimit) {
			uint32_t h = hash4(rd32(ip));
			int32_t cand = table[h];
			table[h] = (int32_t)(ip - base);
			if (cand >= 0 && ip - (base + cand) <= 65535 && rd32(base + cand) == rd32(ip)) {
				const uint8_t *m = base + cand;
				int ml = MINMATCH, lit = (int)(ip - anchor);
				uint8_t *token;
				while (ip + ml < matchlimit && ip[ml] == m[ml]) ml++;
				if (op + lit + lit / 255 + ml / 255 + 16 > oend) return 0;
				token = op++;
				if (lit >= 15) { *token = 15 << 4; op = put_len(op, lit - 15); }
				else *token = (uint8_t)(lit << 4);
				memcpy(op, anchor, (size_t)lit);
				op += lit;
				*op++ = (uint8_t)((ip - m) & 0xff);
				*op++ = (uint8_t)((ip - m) >> 8);
				if (ml - MINMATCH >= 15) { *token |= 15; op = put_len(op, ml - MINMATCH - 15); }
				else *token |= (uint8_t)(ml - MINMATCH);
				ip += ml;
				anchor = ip;
				continue;
			}
			ip++;
		}
	}
	{
		int lit = (int)(end - anchor);
		if (op + lit + lit / 255 + 2 > oend) return 0;
		if (lit >= 15) { *op++ = 15 << 4; op = put_len(op, lit - 15); }
		else *op++ = (uint8_t)(lit << 4);
		memcpy(op, anchor, (size_t)lit);
		op += lit;
	}
	return (int)(op - dst);
}
