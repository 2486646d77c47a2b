// tile_fixup.cpp -- DEVELOPMENT AID (not product, not test): what the fix-up behind the tile kernels (k_parse_fix_lazy) has to do
// when tiles of 16384 positions are parsed from a GUESSED clean state -- how often the guess is wrong, how many 1024-position
// rounds are parsed again before the state falls in step with the tile's own parse, and how many of the positions visited on
// the way the tile never searched (each one a chain walk in global memory by a single lane).  Same functions as the kernels
// (b200z_core.cuh).  Also for warm-ups: the tile starts its speculative parse W positions early (W = 0 is the kernel as it is).
// build: g++ -O2 -std=c++17 -I sharpziplib_b200/csrc -I include -o /tmp/tile_fixup tools/tile_fixup.cpp
// run:   /tmp/tile_fixup <file of concatenated buffers> <buffer size> <level>
#include "b200z_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b200z;

int main(int argc, char **argv) {
	if (argc < 4) return 1;
	const uint32_t bs = (uint32_t)atoi(argv[2]);
	const LevelParams lp = level_params(atoi(argv[3]));
	const uint32_t kTileP = 16384, kRoundP = 1024;
	std::vector<uint8_t> buf(bs + 16);
	for (uint32_t W : {0u, 32u, 64u, 128u, 256u}) {
		FILE *f = fopen(argv[1], "rb");
		if (!f) return 1;
		uint64_t bounds = 0, wrong = 0, rounds = 0, lazy = 0, lazy_cands = 0, warm_steps = 0, streams = 0;
		while (fread(buf.data(), 1, bs, f) == bs) {
			streams++;
			const uint32_t n = bs;
			std::vector<uint16_t> link(n, 0);
			{
				std::vector<int64_t> head(32768, -1);
				for (uint32_t p = 0; p + 2 < n; p++) {
					const uint32_t h = hash3(buf[p], buf[p + 1], buf[p + 2]);
					if (head[h] >= 0 && p - head[h] <= (uint32_t)kMaxDist) link[p] = (uint16_t)(p - head[h]);
					head[h] = p;
				}
			}
			std::vector<uint32_t> A(n), B(n);
			for (uint32_t p = 0; p < n; p++) match_search(buf.data(), link.data(), 0u, p, n, lp, A[p], B[p]);
			auto bytef = [&](uint32_t q) { return (uint32_t)buf[q]; };
			auto slowf = [&](uint32_t p, uint32_t m0, uint32_t budget) { return match_search_above(buf.data(), link.data(), p, n, m0, budget, 0u); };
			// a parse from state c up to position `to`; marks the loop tops in vis (if given)
			auto run = [&](ParseCarry c, uint32_t to, std::vector<char> *vis, uint64_t *steps) {
				auto tabf = [&](uint32_t p, uint32_t &a, uint32_t &b) { a = A[p]; b = B[p]; if (vis) (*vis)[p] = 1; };
				while (c.st.p < to) {
					c.last_top = c.st.p;
					uint32_t s2;
					parse_step(c.st, n, lp, 0, tabf, bytef, slowf, s2);
					if (steps) ++*steps;
				}
				return c;
			};
			auto clean = [&](uint32_t p) { ParseCarry c; parse_init(c.st); c.st.p = p; c.last_top = p; return c; };
			// the tiles' own parses (what the tile kernel converges to): visited positions and the state at every round end
			std::vector<char> S(n + 1, 0);
			std::vector<ParseCarry> spec_end((n + kRoundP - 1) / kRoundP);
			for (uint32_t t0 = 0; t0 < n; t0 += kTileP) {
				const uint32_t t1 = n - t0 > kTileP ? t0 + kTileP : n;
				ParseCarry c = clean(t0);
				if (t0 && W) { // warm-up: parse from t0 - W (clean), keep the state reached at the first loop top >= t0
					c = run(clean(t0 - W), t0, nullptr, &warm_steps);
				}
				for (uint32_t r = t0; r < t1; r += kRoundP) {
					c = run(c, r + kRoundP < t1 ? r + kRoundP : t1, &S, nullptr);
					spec_end[r / kRoundP] = c;
				}
			}
			// the truth, tile by tile, and the fix-up
			ParseCarry truth = clean(0);
			for (uint32_t t0 = 0; t0 < n; t0 += kTileP) {
				const uint32_t t1 = n - t0 > kTileP ? t0 + kTileP : n;
				if (t0) {
					bounds++;
					ParseCarry guess = clean(t0);
					if (W) guess = run(clean(t0 - W), t0, nullptr, nullptr);
					else guess.last_top = truth.last_top; // as k_parse_fix compares: irrelevant, the chunk processes at least one loop top
					if (!carry_equal(truth, guess)) {
						wrong++;
						ParseCarry c = truth;
						for (uint32_t r = t0; r < t1; r += kRoundP) {
							std::vector<char> vis(n + 1, 0);
							c = run(c, r + kRoundP < t1 ? r + kRoundP : t1, &vis, nullptr);
							rounds++;
							for (uint32_t p = r; p < r + kRoundP && p < n; p++)
								if (vis[p] && !S[p] && link[p]) {
									lazy++;
									uint32_t d = link[p], cnt = 0, dist = d;
									for (;;) { // hops of the walk (upper bound of its candidates)
										++cnt;
										if (cnt == (uint32_t)lp.chain) break;
										const uint32_t l2 = link[p - dist];
										if (!l2 || dist + l2 >= (uint32_t)kMaxDist) break;
										dist += l2;
									}
									lazy_cands += cnt;
								}
							if (carry_equal(c, spec_end[r / kRoundP])) break;
						}
					}
				}
				truth = run(truth, t1, nullptr, nullptr);
			}
		}
		fclose(f);
		printf("warm-up %3u: %llu tile boundaries, guess wrong at %.1f%%, %.2f rounds parsed again per wrong guess, %.2f lazy searches per wrong guess "
		       "(%.1f chain hops each at most); per stream: %.1f wrong guesses, %.1f rounds, %.1f lazy searches; warm-up steps per tile %.1f\n",
		       W, (unsigned long long)bounds, 100.0 * wrong / (bounds ? bounds : 1), (double)rounds / (wrong ? wrong : 1),
		       (double)lazy / (wrong ? wrong : 1), (double)lazy_cands / (lazy ? lazy : 1), (double)wrong / streams, (double)rounds / streams,
		       (double)lazy / streams, (double)warm_steps / (bounds ? bounds : 1));
	}
	return 0;
}
