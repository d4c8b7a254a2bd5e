// xxh_chain_probe.cu -- how fast can ONE XXH32 accumulator chain run on an sm_100a lane?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xxh_chain_probe xxh_chain_probe.cu && ./xxh_chain_probe
//
// The round is acc' = rotl(acc + x * P2, 13) * P1 (lib/lz4ada.adb:982-985): three dependent operations when written
// that way (IMAD, SHF, IMAD).  With s = acc + x * P2 as the carried value it is s' = rotl(s, 13) * P1 + x' * P2:
// two dependent operations (SHF, IMAD), the product x' * P2 is off the chain.  Third form: rotl(s, 13) =
// lo(s * 2^13) + hi(s * 2^13), so s' = s * (P1 << 13) + hi(s * 2^13) * P1 + c: IMAD.HI then IMAD, both on the
// multiplier pipe (no pipe crossing).  Prints cycles per round for the three forms; development aid, not product.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u;
__constant__ uint32_t k13 = 8192u;

template <int MODE> __global__ void chain(const uint32_t *xs, uint32_t rounds, uint32_t *out, long long *cycles)
{
	__shared__ uint32_t sx[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) sx[i] = xs[i];
	__syncthreads();
	uint32_t acc = 0x12345678u + threadIdx.x;
	const long long t0 = clock64();
	for (uint32_t r = 0; r < rounds; r += 16) {
		uint32_t x[16];
#pragma unroll
		for (int j = 0; j < 16; j++) x[j] = sx[((r + j) * 4 + (threadIdx.x & 3)) & 255];
		if (MODE == 0) {
#pragma unroll
			for (int j = 0; j < 16; j++) acc = __funnelshift_l(acc + x[j] * P2, acc + x[j] * P2, 13) * P1;
		} else if (MODE == 1) {
			uint32_t s = acc + x[0] * P2;
#pragma unroll
			for (int j = 1; j < 16; j++) {
				uint32_t c;
				const uint32_t r = __funnelshift_l(s, s, 13);
				asm("mul.lo.u32 %0, %1, %2;" : "=r"(c) : "r"(x[j]), "r"(P2));
				asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(r), "r"(P1), "r"(c));
			}
			acc = __funnelshift_l(s, s, 13) * P1;
		} else {
			uint32_t s = acc + x[0] * P2;
			const uint32_t k = k13;
#pragma unroll
			for (int j = 1; j < 16; j++) {
				// (inline PTX: the compiler otherwise re-associates the sum and puts x * P2 back on the chain)
				uint32_t c, hi, a;
				asm("mul.lo.u32 %0, %1, %2;" : "=r"(c) : "r"(x[j]), "r"(P2));
				asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(s), "r"(k));
				asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(s), "r"(P1 << 13), "r"(c));
				asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(hi), "r"(P1), "r"(a));
			}
			acc = __funnelshift_l(s, s, 13) * P1;
		}
	}
	const long long t1 = clock64();
	out[threadIdx.x] = acc;
	if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main()
{
	uint32_t h[256], *dx, *dout;
	long long *dc, c[3];
	uint32_t res[3];
	for (int i = 0; i < 256; i++) h[i] = 0x9e3779b9u * (i + 1);
	cudaMalloc(&dx, sizeof h); cudaMalloc(&dout, 128 * 4); cudaMalloc(&dc, 8);
	cudaMemcpy(dx, h, sizeof h, cudaMemcpyHostToDevice);
	const uint32_t rounds = 1u << 20;
	for (int rep = 0; rep < 2; rep++) {
		chain<0><<<1, 32>>>(dx, rounds, dout, dc); cudaMemcpy(&c[0], dc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&res[0], dout, 4, cudaMemcpyDeviceToHost);
		chain<1><<<1, 32>>>(dx, rounds, dout, dc); cudaMemcpy(&c[1], dc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&res[1], dout, 4, cudaMemcpyDeviceToHost);
		chain<2><<<1, 32>>>(dx, rounds, dout, dc); cudaMemcpy(&c[2], dc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&res[2], dout, 4, cudaMemcpyDeviceToHost);
	}
	if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
	printf("{\"rounds\": %u, \"cycles_per_round\": {\"imad_shf_imad\": %.2f, \"shf_imad\": %.2f, \"imadhi_imad\": %.2f}, \"same_result\": %s}\n", rounds,
	       double(c[0]) / rounds, double(c[1]) / rounds, double(c[2]) / rounds, (res[0] == res[1] && res[1] == res[2]) ? "true" : "false");
	return 0;
}
