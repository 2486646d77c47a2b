// One-off measurement behind DESIGN.md 4: how many bits a decoder started at a wrong bit offset inside a dynamic block needs
// until it is on a true symbol boundary again.  g++ -O2 -o sync tools/huffman_sync_stats.cpp; ./sync stream.deflate [sub-chunk bits]
// Text (C2): median 115, p90 401, p99 840 bits; the high-entropy class of the mix: median 2582, p90 10971.
// measures the self-synchronisation distance of speculative decoding inside the dynamic blocks of a raw deflate stream
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
using namespace std;
static vector<uint8_t> d;
static inline uint32_t bitat(uint64_t p) { return p < 8ull * d.size() ? (d[p >> 3] >> (p & 7)) & 1 : 0; }
static inline uint32_t bits(uint64_t p, int n) { uint32_t v = 0; for (int i = 0; i < n; i++) v |= bitat(p + i) << i; return v; }
static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
static const uint16_t cplens[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t cplext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint8_t cpdext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
struct Code { // canonical decode, bit by bit
	int count[16], first[16], offs[16]; vector<int> sorted;
	void build(const uint8_t *lens, int n) {
		memset(count, 0, sizeof count);
		for (int i = 0; i < n; i++) count[lens[i]]++;
		count[0] = 0;
		int code = 0, off = 0;
		for (int L = 1; L <= 15; L++) { first[L] = code; offs[L] = off; off += count[L]; code = (code + count[L]) << 1; }
		sorted.assign(off, 0);
		int nx[16]; for (int L = 1; L <= 15; L++) nx[L] = offs[L];
		for (int i = 0; i < n; i++) if (lens[i]) sorted[nx[lens[i]]++] = i;
	}
	int decode(uint64_t &p) const { // -1 invalid
		int c = 0;
		for (int L = 1; L <= 15; L++) { c = (c << 1) | bitat(p + L - 1); int idx = c - first[L]; if (idx >= 0 && idx < count[L]) { p += L; return sorted[offs[L] + idx]; } }
		return -1;
	}
};
// decodes one symbol at p; returns 0 ok, 1 eob, -1 error
static int step(const Code &lit, const Code &dst, uint64_t &p) {
	int s = lit.decode(p);
	if (s < 0) return -1;
	if (s < 256) return 0;
	if (s == 256) return 1;
	if (s > 285) return -1;
	p += cplext[s - 257];
	int ds = dst.decode(p);
	if (ds < 0 || ds > 29) return -1;
	p += cpdext[ds];
	return 0;
}
int main(int argc, char **argv) {
	int S = argc > 2 ? atoi(argv[2]) : 512;
	FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); d.resize(n); fread(d.data(), 1, n, f); fclose(f);
	uint64_t p = 0; vector<long> dists; long fails = 0, nsym_total = 0;
	for (;;) {
		int last = bits(p, 1), type = bits(p + 1, 2); p += 3;
		if (type != 2) { fprintf(stderr, "non-dynamic block type %d\n", type); break; }
		int nlit = bits(p, 5) + 257, ndist = bits(p + 5, 5) + 1, nmeta = bits(p + 10, 4) + 4; p += 14;
		uint8_t ml[19] = {0}; for (int i = 0; i < nmeta; i++) { ml[order[i]] = bits(p, 3); p += 3; }
		Code mc; mc.build(ml, 19);
		uint8_t lens[320]; int idx = 0;
		while (idx < nlit + ndist) { int s = mc.decode(p); if (s < 16) lens[idx++] = s; else { int rep; uint8_t v = 0; if (s == 16) { v = lens[idx - 1]; rep = 3 + bits(p, 2); p += 2; } else if (s == 17) { rep = 3 + bits(p, 3); p += 3; } else { rep = 11 + bits(p, 7); p += 7; } while (rep--) lens[idx++] = v; } }
		Code lit, dst; lit.build(lens, nlit); dst.build(lens + nlit, ndist);
		uint64_t b0 = p; vector<uint64_t> starts;
		for (;;) { starts.push_back(p); int r = step(lit, dst, p); if (r) break; }
		uint64_t b1 = p; nsym_total += starts.size();
		// speculative starts
		for (uint64_t q = b0 + S; q + 4 * S < b1; q += S) {
			uint64_t x = q; long steps = 0; bool merged = false;
			while (x < b1) {
				if (binary_search(starts.begin(), starts.end(), x)) { merged = true; break; }
				int r = step(lit, dst, x); steps++;
				if (r) break;
			}
			if (merged) dists.push_back((long)(x - q)); else fails++;
		}
		if (last) break;
	}
	sort(dists.begin(), dists.end());
	size_t m = dists.size();
	printf("%s: symbols %ld, speculative starts %zu (+%ld never merged): sync distance bits  p10 %ld  p50 %ld  p90 %ld  p99 %ld  max %ld  mean %.0f\n", argv[1], nsym_total, m, fails,
	       dists[m / 10], dists[m / 2], dists[m * 9 / 10], dists[m * 99 / 100], dists[m - 1], [&] { double s = 0; for (long v : dists) s += v; return s / m; }());
}
