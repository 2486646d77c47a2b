// tile_passes.cpp -- DEVELOPMENT AID (not product, not test): pass structure of k_tile_parse2 ("guess, batch, re-parse",
// sharpziplib_b200/csrc/experimental/k_tile_parse.cuh) on the CPU, with the kernels' own functions (b200z_core.cuh), for
// two variations.  (1) From which pass on a segment that meets an entry nobody has computed searches it itself, exactly and at
// once (as k_tile_parse does), instead of taking a proxy and waiting for the next batch: the tail passes hold a percent of the
// positions and each costs a CTA-wide barrier round plus a lock-step parse loop.  (2) A dirty segment ends its re-parse at the
// first loop top its previous parse also had in the same simple state, behind that parse's last proxy, and takes the rest
// (exit, symbol count) from it; the tool checks that this is exact.
// Findings on 32 x 256 KiB of the bench mix, level 6: (1) does not pay -- 5.69 -> 4.62 passes from pass 2 on for +0.25 warp
// candidate-steps per position in lock step; the pass count is set by the exit -> entry ripple, not by the proxies.  (2) is
// exact and cuts 188 k re-parses short but the lock-step parse iterations only go from 0.080 to 0.066 per position: a warp's
// iteration count is its slowest lane's, and some lane of a dirty warp usually re-parses in full.
// build: g++ -O2 -std=c++17 -I sharpziplib_b200/csrc -o /tmp/tile_passes tools/tile_passes.cpp
// run:   /tmp/tile_passes <file of concatenated buffers> <buffer size> <level>
#include "b200z_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b200z;

static void links(const uint8_t *data, uint32_t n, std::vector<uint16_t> &link) {
	link.assign(n, 0);
	std::vector<int64_t> head(32768, -1);
	for (uint32_t p = 0; p + 2 < n; p++) {
		const uint32_t h = hash3(data[p], data[p + 1], data[p + 2]);
		if (head[h] >= 0 && p - head[h] <= (uint32_t)kMaxDist) link[p] = (uint16_t)(p - head[h]);
		head[h] = p;
	}
}

// candidates the full-budget walk from threshold 2 looks at (what a batch thread or an on-demand search does)
static uint32_t walk_cands(const uint8_t *data, const uint16_t *link, uint32_t p, uint32_t n, const LevelParams &lp) {
	const uint32_t la = n - p;
	if (la < 3) return 0;
	uint32_t d = link[p];
	if (!d) return 0;
	const uint32_t maxlen = la < 258u ? la : 258u, nice = la < (uint32_t)lp.nice ? la : (uint32_t)lp.nice;
	uint32_t m = 2, dist = d, cnt = 0;
	for (;;) {
		++cnt;
		const uint8_t *c = data + p - dist, *s = data + p;
		if (c[m] == s[m] && c[m - 1] == s[m - 1] && c[0] == s[0] && c[1] == s[1]) {
			uint32_t l = 2;
			while (l < maxlen && c[l] == s[l]) ++l;
			if (l > m) {
				m = l;
				if (m >= nice) break;
			}
		}
		if (cnt == (uint32_t)lp.chain) break;
		const uint32_t l2 = link[p - dist];
		if (!l2) break;
		dist += l2;
		if (dist >= (uint32_t)kMaxDist) break;
	}
	return cnt;
}

int main(int argc, char **argv) {
	if (argc < 4) return 1;
	const uint32_t bs = (uint32_t)atoi(argv[2]);
	const LevelParams lp = level_params(atoi(argv[3]));
	const uint32_t kSeg = 32, kFTile = 16384, kFThreads = 512;
	std::vector<uint8_t> buf(bs + 16);
	for (int cfg = 0; cfg < 5; cfg++) {
		// from: first pass whose misses are searched on demand (99 = never: k_tile_parse2 as it is); early: a dirty segment stops
		// its re-parse at the first loop top its previous parse also had in the same simple state (matchLen < MIN_MATCH: the state
		// is then (position, prevAvailable)) behind that parse's last proxy, and takes the rest from it
		const int from = cfg == 1 ? 3 : cfg == 2 ? 2 : cfg == 3 ? 1 : 99;
		const bool early = cfg == 4;
		FILE *f = fopen(argv[1], "rb");
		if (!f) return 1;
		uint64_t tiles = 0, passes_total = 0, pos = 0, batch_cands = 0, batch_req = 0, parse_iters = 0, od_searches = 0, od_cands = 0,
		         od_warp_steps = 0, barrier_rounds = 0, early_cuts = 0, bad = 0;
		int max_pass = 0;
		LevelParams l1 = lp;
		l1.chain = 1;
		while (fread(buf.data(), 1, bs, f) == bs) {
			const uint32_t n = bs;
			std::vector<uint16_t> link;
			links(buf.data(), n, link);
			auto bytef = [&](uint32_t q) { return (uint32_t)buf[q]; };
			auto slowf = [&](uint32_t p, uint32_t m0, uint32_t budget) { return match_search_above(buf.data(), link.data(), p, n, m0, budget, 0u); };
			for (uint32_t t0 = 0; t0 < n; t0 += kFTile) {
				tiles++;
				pos += (n - t0 < kFTile) ? n - t0 : kFTile;
				std::vector<uint32_t> exA(kFTile, 0xFFFFFFFFu), exB(kFTile, 0);
				std::vector<char> asked(kFTile, 0);
				std::vector<ParseCarry> entry(kFThreads), ex(kFThreads), ex_prev(kFThreads);
				std::vector<uint32_t> lim(kFThreads);
				std::vector<char> dirty(kFThreads, 1), used_proxy(kFThreads, 0);
				std::vector<uint32_t> m0(kFThreads, 0), m1(kFThreads, 0), em(kFThreads, 0), cntv(kFThreads, 0); // simple tops (prevAvail 0 / 1), emitting tops, symbols
				std::vector<int64_t> lastpx(kFThreads, -1); // last position whose proxy the segment's current parse used
				std::vector<char> have(kFThreads, 0);
				auto clean = [&](uint32_t p) { ParseCarry c; parse_init(c.st); c.st.p = p; c.last_top = p; return c; };
				for (uint32_t t = 0; t < kFThreads; t++) {
					const uint32_t seg0 = t0 + t * kSeg;
					lim[t] = seg0 + kSeg < n ? seg0 + kSeg : n;
					entry[t] = clean(seg0);
					ex[t] = entry[t];
				}
				int pass = 0;
				for (;; pass++) {
					std::vector<uint32_t> req;
					for (uint32_t w = 0; w < kFThreads / 32; w++) {
						// lock-step: iteration k of the warp costs what its slowest lane does in its k-th parse step
						std::vector<std::vector<uint32_t>> step_cands(32);
						uint32_t longest = 0;
						for (uint32_t l = 0; l < 32; l++) {
							const uint32_t t = w * 32 + l;
							if (!dirty[t]) continue;
							ex_prev[t] = ex[t];
							ex[t] = entry[t];
							used_proxy[t] = 0;
							uint32_t k = 0, cur = 0;
							auto tabf = [&](uint32_t p, uint32_t &a, uint32_t &b) {
								const uint32_t i = p - t0;
								if (exA[i] != 0xFFFFFFFFu) { a = exA[i]; b = exB[i]; return; }
								if (!link[p]) { exA[i] = 0; exB[i] = 0; a = b = 0; return; }
								if (pass >= from) { // on demand, exact, at once
									match_search(buf.data(), link.data(), 0u, p, n, lp, exA[i], exB[i]);
									a = exA[i]; b = exB[i];
									const uint32_t c = walk_cands(buf.data(), link.data(), p, n, lp);
									od_searches++;
									od_cands += c;
									cur += c;
									return;
								}
								match_search(buf.data(), link.data(), 0u, p, n, l1, a, b);
								used_proxy[t] = 1;
								if (!asked[i]) { asked[i] = 1; req.push_back(p); }
							};
							const ParseCarry old_ex = ex_prev[t];
							const uint32_t o0 = m0[t], o1 = m1[t], oe = em[t];
							const int64_t olp = lastpx[t];
							uint32_t n0 = 0, n1 = 0, ne = 0, cn = 0;
							int64_t nlp = -1;
							const uint32_t seg0 = t0 + t * kSeg;
							bool cut = false;
							while (ex[t].st.p < lim[t]) {
								const uint32_t i = ex[t].st.p - seg0;
								const bool simple = ex[t].st.mlen < (uint32_t)kMinMatch;
								if (early && have[t] && simple && (((ex[t].st.prevAvail ? o1 : o0) >> i) & 1u) && (int64_t)ex[t].st.p > olp) {
									// the rest is the previous parse's
									const uint32_t keep = ~((1u << i) - 1u);
									n0 |= o0 & keep;
									n1 |= o1 & keep;
									ne |= oe & keep;
									cn += (uint32_t)__builtin_popcount(oe & keep);
									ex[t] = old_ex;
									cut = true;
									early_cuts++;
									break;
								}
								if (simple) (ex[t].st.prevAvail ? n1 : n0) |= 1u << i;
								ex[t].last_top = ex[t].st.p;
								uint32_t s2;
								cur = 0;
								const char up0 = used_proxy[t];
								used_proxy[t] = 0;
								if (parse_step(ex[t].st, n, lp, 0, tabf, bytef, slowf, s2)) { ne |= 1u << i; cn++; }
								if (used_proxy[t]) nlp = (int64_t)seg0 + i + 1; // (a step consults at most positions i and i + 1: be conservative)
								used_proxy[t] |= up0;
								step_cands[l].push_back(cur);
								k++;
							}
							(void)cut;
							m0[t] = n0; m1[t] = n1; em[t] = ne; cntv[t] = cn; lastpx[t] = nlp; have[t] = 1;
							if (k > longest) longest = k;
						}
						parse_iters += longest;
						for (uint32_t k = 0; k < longest; k++) {
							uint32_t mx = 0;
							for (uint32_t l = 0; l < 32; l++)
								if (k < step_cands[l].size() && step_cands[l][k] > mx) mx = step_cands[l][k];
							od_warp_steps += mx;
						}
					}
					for (uint32_t p : req) {
						match_search(buf.data(), link.data(), 0u, p, n, lp, exA[p - t0], exB[p - t0]);
						batch_cands += walk_cands(buf.data(), link.data(), p, n, lp);
					}
					batch_req += req.size();
					barrier_rounds += 3 + (req.size() >= 4 * kFThreads ? 4 : 0);
					bool again = false;
					std::vector<ParseCarry> prev(ex);
					for (uint32_t t = 0; t < kFThreads; t++) {
						bool d = used_proxy[t] != 0;
						if (t > 0 && !carry_equal(prev[t - 1], entry[t])) { entry[t] = prev[t - 1]; d = true; }
						dirty[t] = d;
						again |= d;
					}
					if (!again) break;
				}
				if (early) { // every segment's exit and symbol count must be what a full parse from its entry gives with exact entries
					for (uint32_t t = 0; t < kFThreads; t++) {
						ParseCarry c = entry[t];
						uint32_t cn = 0;
						auto tabx = [&](uint32_t p, uint32_t &a, uint32_t &b) { match_search(buf.data(), link.data(), 0u, p, n, lp, a, b); };
						while (c.st.p < lim[t]) {
							c.last_top = c.st.p;
							uint32_t s2;
							cn += (uint32_t)parse_step(c.st, n, lp, 0, tabx, bytef, slowf, s2);
						}
						if (!carry_equal(c, ex[t]) || cn != cntv[t]) bad++;
					}
				}
				passes_total += (uint64_t)pass + 1;
				if (pass + 1 > max_pass) max_pass = pass + 1;
			}
		}
		fclose(f);
		if (early) printf("early exit of re-parses: %llu cuts, %llu segments differ from a full parse (must be 0)\n", (unsigned long long)early_cuts, (unsigned long long)bad);
		printf("on demand from pass %2d: %.2f passes per tile (max %d), %.2f barrier rounds per tile; per position: %.3f lock-step parse iterations, "
		       "batched %.1f%% with %.2f candidates, on demand %.2f%% with %.2f candidates = %.3f warp candidate-steps in lock step\n",
		       from, (double)passes_total / tiles, max_pass, (double)barrier_rounds / tiles, (double)parse_iters / pos, 100.0 * batch_req / pos,
		       (double)batch_cands / pos, 100.0 * od_searches / pos, (double)od_cands / pos, (double)od_warp_steps / pos);
	}
	return 0;
}
