// cuda_emu.cpp -- TEST INFRASTRUCTURE: the block scheduler of cuda_emu.h
#include "cuda_emu.h"

#include <algorithm>
#include <vector>

namespace emu {

Block *g_block = nullptr;
Fiber *g_cur = nullptr;
uint64_t g_events = 0;
size_t g_stack_bytes = 0;

static void fiber_main() {
	g_block->body();
	Fiber *f = g_cur;
	Block &b = *g_block;
	f->done = true;
	g_events++;
	// a thread that has left no longer takes part in barriers (CUDA: exited threads are not waited for)
	Warp &w = b.warps[f->warp];
	w.live--;
	b.live--;
	if (w.live > 0 && w.arrived >= w.live) {
		w.arrived = 0;
		w.gen++;
	}
	if (b.live > 0 && b.arrived >= b.live) {
		b.arrived = 0;
		b.gen++;
	}
	swapcontext(&f->ctx, &b.sched);
}

static std::vector<void *> g_stacks;
static uint8_t *g_smem = nullptr;

void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body) {
	const int nt = (int)(block.x * block.y * block.z);
	if (smem_bytes > kMaxSmem) {
		fprintf(stderr, "cuda_emu: %zu bytes of dynamic shared memory\n", smem_bytes);
		abort();
	}
	if (!g_smem) g_smem = (uint8_t *)aligned_alloc(256, kMaxSmem + 4096);
	if (!g_stack_bytes) {
		const char *e = getenv("B200Z_EMU_STACK_KB");
		g_stack_bytes = (size_t)(e ? atoi(e) : 1024) << 10;
	}
	while ((int)g_stacks.size() < nt) {
		void *s = mmap(nullptr, g_stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (s == MAP_FAILED) abort();
		g_stacks.push_back(s);
	}
	// B200Z_EMU_ORDER: the order in which the threads of a block are resumed in a scheduling round, and the order of the blocks
	// of a grid.  "fwd" (default): ascending.  "rev": descending.  "rand:<seed>": a fresh pseudo-random permutation per round,
	// blocks in a random order.  A kernel whose result depends on the order has a race (a read of another thread's write
	// without a barrier in between, an assumption about which block runs first): run the tests under all three.
	static int order_mode = -1;
	static uint64_t rng = 0;
	if (order_mode < 0) {
		const char *e = getenv("B200Z_EMU_ORDER");
		order_mode = 0;
		if (e && !strcmp(e, "rev")) order_mode = 1;
		if (e && !strncmp(e, "rand", 4)) {
			order_mode = 2;
			rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)(e[4] == ':' ? atoll(e + 5) : 1);
		}
	}
	auto next_rand = [&]() {
		rng ^= rng << 13;
		rng ^= rng >> 7;
		rng ^= rng << 17;
		return rng;
	};
	auto permute = [&](std::vector<int> &v) {
		const int n = (int)v.size();
		for (int i = 0; i < n; i++) v[(size_t)i] = order_mode == 1 ? n - 1 - i : i;
		if (order_mode == 2)
			for (int i = n - 1; i > 0; i--) std::swap(v[(size_t)i], v[(size_t)(next_rand() % (uint64_t)(i + 1))]);
	};
	std::vector<int> torder((size_t)nt), border((size_t)(grid.x * grid.y * grid.z));
	permute(border);
	Block b;
	b.bdim = block;
	b.gdim = grid;
	b.body = body;
	b.smem = g_smem;
	for (size_t bi = 0; bi < border.size(); bi++) {
				const unsigned lin = (unsigned)border[bi];
				const unsigned bx = lin % grid.x, by = (lin / grid.x) % grid.y, bz = lin / (grid.x * grid.y);
				b.bid = dim3(bx, by, bz);
				memset(g_smem, 0xCD, smem_bytes + 64); // shared memory is not cleared between blocks
				if (const char *fe = getenv("B200Z_EMU_FILL")) { // < 0: pseudo-random leftovers instead of a constant (static __shared__ arrays are host statics here and keep what the last block left, like the device)
					const int fv = atoi(fe);
					if (fv < 0) {
						static uint64_t st = (uint64_t)(-fv) * 0xD1B54A32D192ED03ull + 7;
						for (size_t i = 0; i < smem_bytes; i++) {
							st = st * 6364136223846793005ull + 1442695040888963407ull;
							g_smem[i] = (uint8_t)(st >> 56);
						}
					}
				}
				memset(g_smem + smem_bytes, 0xEE, 64);
				b.fibers.assign((size_t)nt, Fiber());
				b.warps.assign((size_t)(nt + 31) / 32, Warp());
				b.live = nt;
				b.arrived = 0;
				b.gen = 0;
				b.or_acc = 0;
				b.or_reset_gen = 0xFFFFFFFFu;
				b.cnt_acc = 0;
				b.cnt_reset_gen = 0xFFFFFFFFu;
				g_block = &b;
				for (int t = 0; t < nt; t++) {
					Fiber &f = b.fibers[(size_t)t];
					f.linear = t;
					f.lane = t & 31;
					f.warp = t >> 5;
					f.tid = dim3((unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y));
					b.warps[(size_t)f.warp].live++;
					getcontext(&f.ctx);
					f.ctx.uc_stack.ss_sp = g_stacks[(size_t)t];
					f.ctx.uc_stack.ss_size = g_stack_bytes;
					f.ctx.uc_link = nullptr;
					makecontext(&f.ctx, fiber_main, 0);
				}
				int remaining = nt;
				b.or_reset_gen = 0xFFFFFFFFu;
				while (remaining > 0) {
					const uint64_t before = g_events;
					permute(torder);
					for (int ti = 0; ti < nt; ti++) {
						const int t = torder[(size_t)ti];
						Fiber &f = b.fibers[(size_t)t];
						if (f.done) continue;
						g_cur = &f;
						swapcontext(&b.sched, &f.ctx);
						if (f.done) remaining--;
					}
					if (remaining > 0 && g_events == before) {
						// every thread was resumed and none got past what it waits for
						fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %d of %d threads wait at barriers the others never reach\n",
						        bx, by, bz, remaining, nt);
						abort();
					}
				}
				for (int i = 0; i < 64; i++)
					if (g_smem[smem_bytes + (size_t)i] != 0xEE) {
						fprintf(stderr, "cuda_emu: block (%u,%u,%u) wrote behind its %zu bytes of dynamic shared memory\n", bx, by, bz, smem_bytes);
						abort();
					}
			}
	g_block = nullptr;
	g_cur = nullptr;
}

} // namespace emu
