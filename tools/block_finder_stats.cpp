// One-off measurement behind DESIGN.md 4: survivors of the block finder's stages over every bit offset of a raw deflate stream.
// g++ -O2 -o find tools/block_finder_stats.cpp; ./find stream.deflate
// C2 text stream (2.5 Mbit): 274 428 pass the header fields, 1 102 have a complete code-length code, 10 are block headers (all true).
// counts survivors of each finder stage over every bit offset of a raw deflate stream; also lists true block starts
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
using namespace std;
static vector<uint8_t> d;
static inline uint64_t bits(uint64_t pos, int n) { // n <= 57
	uint64_t v = 0;
	uint64_t by = pos >> 3;
	for (int i = 0; i < 9; i++) v |= (by + i < d.size() ? (uint64_t)d[by + i] : 0ull) << (8 * i) >> 0, (void)0;
	// careful: 9 bytes overflow; do it properly
	unsigned __int128 w = 0;
	for (int i = 0; i < 10; i++) w |= (unsigned __int128)(by + i < d.size() ? d[by + i] : 0) << (8 * i);
	return (uint64_t)(w >> (pos & 7)) & ((n == 64) ? ~0ull : ((1ull << n) - 1));
}
static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
long s1 = 0, s2 = 0, s3 = 0, s3b = 0, s4 = 0;
bool check(uint64_t pos, bool verbose) {
	uint64_t h = bits(pos, 17);
	if ((h & 1) != 0) return false; // BFINAL=0 only
	if (((h >> 1) & 3) != 2) return false;
	int hlit = (h >> 3) & 31, hdist = (h >> 8) & 31, hclen = (h >> 13) & 15;
	if (hlit > 29 || hdist > 29) return false;
	s1++;
	int nmeta = hclen + 4;
	uint8_t ml[19] = {0};
	uint64_t p = pos + 17;
	int kraft = 0, nz = 0;
	for (int i = 0; i < nmeta; i++) { int l = bits(p, 3); p += 3; ml[order[i]] = l; if (l) { kraft += 128 >> l; nz++; } }
	if (kraft != 128) return false; // complete precode only
	s2++;
	// canonical decode table (7 bits)
	int count[8] = {0}, next[8];
	for (int i = 0; i < 19; i++) count[ml[i]]++;
	count[0] = 0;
	int code = 0;
	for (int L = 1; L <= 7; L++) { next[L] = code; code = (code + count[L]) << 1; }
	uint8_t tab_sym[128], tab_len[128];
	memset(tab_len, 0, sizeof tab_len);
	for (int s = 0; s < 19; s++) {
		int L = ml[s]; if (!L) continue;
		int c = next[L]++;
		int rev = 0; for (int b = 0; b < L; b++) if (c >> b & 1) rev |= 1 << (L - 1 - b);
		for (int i = rev; i < 128; i += 1 << L) { tab_sym[i] = s; tab_len[i] = L; }
	}
	int nlit = hlit + 257, ndist = hdist + 1, total = nlit + ndist;
	uint8_t lens[320];
	int idx = 0;
	while (idx < total) {
		int v = bits(p, 7);
		int L = tab_len[v]; if (!L) return false;
		int sym = tab_sym[v]; p += L;
		if (sym < 16) lens[idx++] = sym;
		else {
			int rep; uint8_t val = 0;
			if (sym == 16) { if (idx == 0) return false; val = lens[idx - 1]; rep = 3 + bits(p, 2); p += 2; }
			else if (sym == 17) { rep = 3 + bits(p, 3); p += 3; }
			else { rep = 11 + bits(p, 7); p += 7; }
			if (idx + rep > total) return false;
			while (rep-- > 0) lens[idx++] = val;
		}
	}
	if (p > 8ull * d.size()) return false;
	s3++;
	if (lens[256] == 0) return false;
	s3b++;
	long kl = 0, kd = 0; int nd = 0;
	for (int i = 0; i < nlit; i++) if (lens[i]) kl += 32768 >> lens[i];
	for (int i = 0; i < ndist; i++) if (lens[nlit + i]) { kd += 32768 >> lens[nlit + i]; nd++; }
	if (kl != 32768) return false;
	if (!(kd == 32768 || nd <= 1)) return false;
	s4++;
	if (verbose) printf("  cand at bit %llu (byte %llu) hdr bits %llu\n", (unsigned long long)pos, (unsigned long long)pos / 8, (unsigned long long)(p - pos));
	return true;
}
int main(int argc, char **argv) {
	for (int a = 1; a < argc; a++) {
		FILE *f = fopen(argv[a], "rb");
		fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
		d.resize(n); fread(d.data(), 1, n, f); fclose(f);
		s1 = s2 = s3 = s3b = s4 = 0;
		for (uint64_t pos = 0; pos < 8ull * n; pos++) check(pos, true);
		printf("%s: bits %ld  s1(hdr fields) %ld  s2(precode complete) %ld  s3(lengths decode) %ld s3b(eob) %ld s4(complete codes) %ld\n", argv[a], 8 * n, s1, s2, s3, s3b, s4);
	}
}
