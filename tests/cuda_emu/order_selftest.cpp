// TEST INFRASTRUCTURE: self-test of B200Z_EMU_ORDER.  A deliberately racy kernel (every thread reads its right neighbour's
// slot without a barrier after the write) must give different results under "fwd" and "rev", a correct one (barrier in
// between) the same result under every order; prints "racy=<sum> clean=<sum>".
#include "cuda_emu.h"

static uint32_t *g_out;
static void k_racy(int with_barrier) {
	__shared__ uint32_t s[64];
	const int t = threadIdx.x;
	s[t] = 0;
	__syncthreads();
	s[t] = (uint32_t)t + 1;
	if (with_barrier) __syncthreads();
	g_out[t] = s[(t + 1) & 63];
}

int main() {
	uint32_t out[64];
	g_out = out;
	uint64_t sums[2];
	for (int wb = 0; wb < 2; wb++) {
		emu::launch(dim3(1), dim3(64), 0, [=]() { k_racy(wb); });
		uint64_t s = 0;
		for (int i = 0; i < 64; i++) s += out[i];
		sums[wb] = s;
	}
	printf("racy=%llu clean=%llu\n", (unsigned long long)sums[0], (unsigned long long)sums[1]);
	return 0;
}
