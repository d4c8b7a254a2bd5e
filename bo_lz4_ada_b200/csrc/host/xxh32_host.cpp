// xxh32_host.cpp -- package XXHash32 on the host (lib/lz4ada.ads:311-344, lib/lz4ada.adb:923-1026).
// Serves the 1-byte frame-header checksum and the public hasher API (tool_xxhash32ada's use);
// block and content checksums of the decode path run on the device (kernels.cuh).
#include "common.hpp"

namespace lz4ada {

static const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;

static inline uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

static inline void stripe(lz4ada_xxhash32 *h, const uint8_t *d)
{
	for (int i = 0; i < 4; i++) h->state[i] = rol(h->state[i] + load32(d + 4 * i) * P2, 13) * P1;
}

void Xxh32Host::reset(lz4ada_xxhash32 *h, uint32_t seed)
{
	h->state[0] = seed + P1 + P2;
	h->state[1] = seed + P2;
	h->state[2] = seed;
	h->state[3] = seed - P1;
	h->buffer_size = 0;
	h->total_length = 0;
}

void Xxh32Host::update(lz4ada_xxhash32 *h, const uint8_t *p, size_t n)
{
	h->total_length += n;
	if (h->buffer_size > 0) {   // top up a partial stripe first
		size_t take = 16 - size_t(h->buffer_size);
		if (take > n) take = n;
		memcpy(h->buffer + h->buffer_size, p, take);
		h->buffer_size += int32_t(take);
		p += take;
		n -= take;
		if (h->buffer_size < 16) return;
		stripe(h, h->buffer);
		h->buffer_size = 0;
	}
	for (; n >= 16; p += 16, n -= 16) stripe(h, p);
	if (n) {
		memcpy(h->buffer, p, n);
		h->buffer_size = int32_t(n);
	}
}

uint32_t Xxh32Host::final(const lz4ada_xxhash32 *h)
{
	uint32_t r = uint32_t(h->total_length);
	r += h->total_length >= 16 ? rol(h->state[0], 1) + rol(h->state[1], 7) + rol(h->state[2], 12) +
					     rol(h->state[3], 18)
				   : h->state[2] + P5;
	int i = 0;
	for (; i + 4 <= h->buffer_size; i += 4) r = rol(r + load32(h->buffer + i) * P3, 17) * P4;
	for (; i < h->buffer_size; i++) r = rol(r + uint32_t(h->buffer[i]) * P5, 11) * P1;
	r = (r ^ (r >> 15)) * P2;
	r = (r ^ (r >> 13)) * P3;
	return r ^ (r >> 16);
}

uint32_t Xxh32Host::hash(const uint8_t *p, size_t n)
{
	lz4ada_xxhash32 h;
	reset(&h, 0);
	update(&h, p, n);
	return final(&h);
}

}  // namespace lz4ada
