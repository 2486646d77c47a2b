// b200z_inflate_par.cuh -- the block-parallel inflate pipeline (included by b200z_inflate.cu).
//
// The reference's Inflater (Zip/Compression/Inflater.cs:429-552) walks a stream bit by bit; across a block boundary it
// carries nothing but the bit position and the window (= the output so far).  That is the parallelism used here:
//
//   k_find / k_find3   every bit offset of every stream is tested for "a dynamic block header could start here" (BFINAL = 0,
//                      BTYPE = 2, HLIT / HDIST in range, a complete code-length code, code lengths that decode to exactly
//                      HLIT + HDIST entries with an end-of-block code and complete literal and distance codes).  The first
//                      survivor of every 8192-bit window of a stream is a CANDIDATE.  Candidates are hints, nothing more.
//   k_seglist          the stream starts and the candidates become the SEGMENT list.
//   k_dec1             persistent CTAs take segments off the list and decode them without writing output: block headers
//                      (InflaterDynHeader.cs:42-120), code tables (InflaterHuffmanTree.cs:87-169) and ROUNDS of 128 lanes x
//                      512 bits -- every lane decodes its sub-chunk speculatively, exit -> entry hand-off until stable, the
//                      same scheme as the serial kernel, four warps wide.  A segment ends when the position it has reached
//                      at a block boundary is itself a candidate (someone else decodes on from there: the two decoders are
//                      the same deterministic function of the bit position, so the hand-over is exact), at the final block,
//                      at an error or at the end of the input.  Per round it records every lane's entry position and the
//                      prefix sums of the bytes and back-references the lanes produce.
//   k_chain            one thread per stream follows the segments from the stream start (join -> the candidate's segment
//                      -> ...), gives every segment on the chain its output position, and derives what Inflater reports:
//                      status, TotalOut, RemainingInput (trap T14), the restart point.
//   k_dec2             one CTA per recorded batch of rounds: the final decode pass, every lane from its recorded entry with
//                      its recorded output position -- literals straight into the output buffer, back-references as
//                      (position, length, distance) records in stream order.
//   k_resolve          one CTA per stream walks the output in 16 KiB tiles: every byte of a back-reference points at its
//                      source byte (OutputWindow.Repeat's rule, Streams/OutputWindow.cs:63-92: byte k of a copy comes from
//                      source byte k mod distance), sources in front of the tile are final and read directly, pointers inside
//                      the tile are resolved by pointer jumping in shared memory, the tile is written back.
//
// Anything that does not fit the pipeline's fixed pools (pathological streams of thousands of tiny blocks, outputs beyond
// 4 GiB) is flagged and decoded by the serial kernel k_inflate afterwards; results are identical by construction, the
// serial kernel is the reference point the parallel one is tested against.
#pragma once

namespace b200z {

constexpr int kFWShift = 13; // candidate windows of 8192 bits
constexpr uint32_t kFWMask = (1u << kFWShift) - 1u;
constexpr uint32_t kNoCand = 0xFFFFFFFFu;
constexpr int kP1Threads = 128;                              // lanes of a round
constexpr int kP1RoundWords = kP1Threads * kSubBits / 32;   // 2048 words of input per round
constexpr int kP1InWords = kP1RoundWords + 8;
constexpr int kP1InSlots = kP1InWords + kP1InWords / 16 + 1;
constexpr int kRoundBatch = 4;                               // rounds are allocated four at a time, a batch never mixes blocks
constexpr int kFindTileWords = 2048;
constexpr int kResTile = 16384;                              // k_resolve: bytes per tile
constexpr int kResThreads = 1024;

enum { SEG_JOIN = 0, SEG_FINAL = 1, SEG_STOP = 2, SEG_POOL = 3 };
enum { RND_NONE = 0, RND_HUFF = 1, RND_STORED = 2 };

struct FTile { // k_find: 2048 words of one stream
	uint32_t stream, word0, nwords, pad;
};

struct __align__(16) PSeg {
	uint64_t start_bit, end_bit, rs_bit;
	uint64_t out_bytes, rs_out; // relative to the segment's first output byte
	uint64_t out_base;          // k_chain: output position of the segment
	uint32_t stream, n_match, match_base;
	int32_t end_kind, status;
	uint32_t valid;
};

struct __align__(16) PRound {
	uint32_t seg, kind, hdr, w0;            // hdr: index of the block's PBlockHdr; w0: first staged word (stored: source byte offset)
	uint32_t r0, end_rel, lastlane, nbytes; // stored: nbytes = LEN
	uint64_t out_rel;                       // relative to the segment
	uint32_t match_rel, pad;
	uint32_t entry[kP1Threads];
	uint32_t opre[kP1Threads]; // exclusive prefix sums over the lanes: output bytes / back-references
	uint32_t mpre[kP1Threads];
};

struct __align__(16) PBlockHdr {
	uint16_t nlit, ndist, is_static, pad;
	uint8_t lens[320];
};

struct PCounters {
	uint32_t fs_count, nseg, seg_next, round_top, hdr_top, pad[3];
};

// ---------------------------------------------------------------------------------------------------------
// finder
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_stream_word(const uint32_t *gwords, uint32_t wi, uint32_t nwords, uint32_t nbytes) {
	uint32_t w = 0;
	if (wi < nwords) {
		w = __ldg(gwords + wi);
		if (wi == nwords - 1 && (nbytes & 3)) w &= (1u << (8 * (nbytes & 3))) - 1u;
	}
	return w;
}

// Stages 1 and 2.  32 bit offsets at a time: the fixed header fields as bit-parallel masks over funnel-shifted words, then
// the Kraft sum of the code-length code (a complete code: sum of 2^(7-len) = 128) through a table of four lengths at a time.
__global__ void __launch_bounds__(256)
    k_find(const uint8_t *__restrict__ in, const int64_t *__restrict__ in_off, const int64_t *__restrict__ in_len,
           const FTile *__restrict__ tiles, const uint32_t *__restrict__ start_bit, const int32_t *__restrict__ pre,
           unsigned long long *__restrict__ fs_list, PCounters *__restrict__ ctr, uint32_t fs_cap) {
	__shared__ uint8_t lut[4096];
	for (int i = threadIdx.x; i < 4096; i += 256) {
		uint32_t s = 0;
		for (int j = 0; j < 4; j++) {
			const uint32_t l = ((uint32_t)i >> (3 * j)) & 7u;
			if (l) s += 128u >> l;
		}
		lut[i] = (uint8_t)(s > 255u ? 255u : s);
	}
	__syncthreads();
	const FTile t = tiles[blockIdx.x];
	if (pre && pre[t.stream] != B200Z_OK) return;
	const uint32_t *gwords = reinterpret_cast<const uint32_t *>(in + in_off[t.stream]);
	const uint32_t nbytes = (uint32_t)in_len[t.stream];
	const uint32_t nwords = (nbytes + 3) >> 2;
	const uint64_t sb = start_bit ? start_bit[t.stream] : 0u;
	const uint64_t total_bits = 8ull * nbytes;
	for (uint32_t k = threadIdx.x; k < t.nwords; k += 256) {
		const uint32_t wi = t.word0 + k;
		const uint32_t w0 = ld_stream_word(gwords, wi, nwords, nbytes), w1 = ld_stream_word(gwords, wi + 1, nwords, nbytes);
#define B200Z_S(kk) __funnelshift_r(w0, w1, kk)
		uint32_t mask = ~w0 & ~B200Z_S(1) & B200Z_S(2);                            // BFINAL = 0, BTYPE = 2 (bits 1, 2 = 0, 1)
		if (mask) mask &= ~(B200Z_S(4) & B200Z_S(5) & B200Z_S(6) & B200Z_S(7));    // HLIT <= 29
		if (mask) mask &= ~(B200Z_S(9) & B200Z_S(10) & B200Z_S(11) & B200Z_S(12)); // HDIST <= 29
#undef B200Z_S
		const uint64_t wbit = 32ull * wi;
		if (wbit + 32 <= sb) continue;
		if (wbit <= sb) { // nothing in front of the stream's first block header, and not that header itself (segment 0 has it)
			const uint32_t cut = (uint32_t)(sb - wbit) + 1;
			mask = cut >= 32 ? 0u : (mask >> cut) << cut;
		}
		if (!mask) continue;
		const uint32_t w2 = ld_stream_word(gwords, wi + 2, nwords, nbytes), w3 = ld_stream_word(gwords, wi + 3, nwords, nbytes);
		const uint64_t lo = (uint64_t)w0 | ((uint64_t)w1 << 32), hi = (uint64_t)w2 | ((uint64_t)w3 << 32);
		while (mask) {
			const uint32_t o = (uint32_t)__ffs((int)mask) - 1u;
			mask &= mask - 1u;
			const uint32_t sh = o + 13; // HCLEN and the code-length code's lengths: 4 + 57 bits
			uint64_t v = (lo >> sh) | (hi << (64 - sh));
			const uint32_t nmeta = (uint32_t)(v & 15u) + 4u;
			v >>= 4;
			v &= (1ull << (3 * nmeta)) - 1ull;
			const uint32_t sum = (uint32_t)lut[v & 4095u] + lut[(v >> 12) & 4095u] + lut[(v >> 24) & 4095u] + lut[(v >> 36) & 4095u] +
			                     lut[(v >> 48) & 4095u];
			if (sum != 128u) continue;
			const uint64_t pos = wbit + o;
			if (pos + 17 + 3 * nmeta > total_bits) continue;
			const uint32_t idx = atomicAdd(&ctr->fs_count, 1u);
			if (idx < fs_cap) fs_list[idx] = ((unsigned long long)t.stream << 40) | pos;
		}
	}
}

// Stage 3, one thread per survivor: the code lengths themselves (InflaterDynHeader.cs:62-117's rules) and the two codes
// they describe.  What passes is entered as its window's candidate (the lowest position wins).
__global__ void __launch_bounds__(128)
    k_find3(const uint8_t *__restrict__ in, const int64_t *__restrict__ in_off, const int64_t *__restrict__ in_len,
            const unsigned long long *__restrict__ fs_list, const PCounters *__restrict__ ctr, uint32_t fs_cap,
            const uint32_t *__restrict__ win_base, uint32_t *__restrict__ cand) {
	uint32_t count = ctr->fs_count;
	if (count > fs_cap) count = fs_cap;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		const unsigned long long e = fs_list[i];
		const uint32_t stream = (uint32_t)(e >> 40);
		const uint64_t pos = e & ((1ull << 40) - 1ull);
		BitReader br;
		br.words = reinterpret_cast<const uint32_t *>(in + in_off[stream]);
		br.nbytes = (uint32_t)in_len[stream];
		br.nwords = (br.nbytes + 3) >> 2;
		br.consumed = pos;
		br.widx = (uint32_t)(pos >> 5);
		br.bb = 0;
		br.bc = 0;
		const uint32_t sk = (uint32_t)(pos & 31);
		if (sk) {
			br.refill();
			br.bb >>= sk;
			br.bc -= sk;
		}
		br.get(3);
		const int nlit = (int)br.get(5) + 257, ndist = (int)br.get(5) + 1, nmeta = (int)br.get(4) + 4;
		uint8_t ml[19];
		for (int k = 0; k < 19; k++) ml[k] = 0;
		for (int k = 0; k < nmeta; k++) ml[c_meta_order[k]] = (uint8_t)br.get(3);
		// 7-bit decode table of the code-length code: sym << 3 | len
		uint8_t tab[128];
		for (int k = 0; k < 128; k++) tab[k] = 0;
		{
			uint32_t cnt[8], nxt[8];
			for (int L = 0; L < 8; L++) cnt[L] = 0;
			for (int s = 0; s < 19; s++) cnt[ml[s]]++;
			cnt[0] = 0;
			uint32_t code = 0;
			for (int L = 1; L <= 7; L++) {
				nxt[L] = code;
				code = (code + cnt[L]) << 1;
			}
			for (int s = 0; s < 19; s++) {
				const int L = ml[s];
				if (!L) continue;
				const uint32_t c = nxt[L]++;
				const uint32_t rev = __brev(c) >> (32 - L);
				for (uint32_t k = rev; k < 128u; k += (1u << L)) tab[k] = (uint8_t)((s << 3) | L);
			}
		}
		const int total = nlit + ndist;
		int idx = 0, prev = 0, nd = 0, eob_len = 0;
		uint32_t kl = 0, kd = 0;
		bool ok = true;
		while (ok && idx < total) {
			br.refill();
			const uint32_t te = tab[br.peek(7)];
			if (!te) { ok = false; break; }
			br.drop((int)(te & 7u));
			const int sym = (int)(te >> 3);
			int rep = 1, val = sym;
			if (sym == 16) {
				if (idx == 0) { ok = false; break; }
				val = prev;
				rep = 3 + (int)br.get(2);
			} else if (sym == 17) {
				val = 0;
				rep = 3 + (int)br.get(3);
			} else if (sym == 18) {
				val = 0;
				rep = 11 + (int)br.get(7);
			}
			if (idx + rep > total) { ok = false; break; }
			for (int r = 0; r < rep; r++, idx++) {
				if (idx < nlit) {
					if (val) kl += 32768u >> val;
					if (idx == 256) eob_len = val;
				} else if (val) {
					kd += 32768u >> val;
					nd++;
				}
			}
			prev = val;
			if (br.overrun()) ok = false;
		}
		if (!ok || br.overrun() || eob_len == 0 || kl != 32768u || !(kd == 32768u || nd <= 1)) continue;
		atomicMin(&cand[win_base[stream] + (uint32_t)(pos >> kFWShift)], (uint32_t)pos & kFWMask);
	}
}

// the segment list: slots 0 .. n-1 are the stream starts, slot n + g is window g's candidate
__global__ void k_seglist(int n, uint32_t nwin_total, const uint32_t *__restrict__ win_base, const uint32_t *__restrict__ win_stream,
                          const uint32_t *__restrict__ cand, const uint32_t *__restrict__ start_bit, const int32_t *__restrict__ pre,
                          PSeg *__restrict__ segs, uint32_t *__restrict__ seg_list, PCounters *__restrict__ ctr) {
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= (uint32_t)n + nwin_total) return;
	uint32_t stream;
	uint64_t sbit;
	if (slot < (uint32_t)n) {
		stream = slot;
		if (pre && pre[stream] != B200Z_OK) return;
		sbit = start_bit ? start_bit[stream] : 0u;
	} else {
		const uint32_t g = slot - (uint32_t)n;
		const uint32_t c = cand[g];
		if (c == kNoCand) return;
		stream = win_stream[g];
		sbit = ((uint64_t)(g - win_base[stream]) << kFWShift) + c;
	}
	PSeg &s = segs[slot];
	s.stream = stream;
	s.start_bit = sbit;
	s.valid = 0;
	s.end_kind = SEG_STOP;
	s.status = B200Z_E_INTERNAL;
	seg_list[atomicAdd(&ctr->nseg, 1u)] = slot;
}

// ---------------------------------------------------------------------------------------------------------
// code tables, built by the whole CTA (the same tables build_table() makes serially)
// ---------------------------------------------------------------------------------------------------------
struct TabScratch {
	uint32_t cnt[2][16], nxt[2][16], run[2][16];
	int err;
};

// lens[0 .. nlit) literal/length code lengths, lens[nlit .. nlit + ndist) distance code lengths.  All threads of the CTA
// (>= 64) call it; returns 0 or a detail code, the same for every thread.
__device__ int build_tables_cta(InfShared &sh, TabScratch &ts, int nlit, int ndist) {
	const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < (1 << kLitRoot); i += nthr) sh.lit[i] = 0;
	for (int i = tid; i < (1 << kDistRoot); i += nthr) sh.dist[i] = 0;
	if (tid < 32) {
		ts.cnt[tid >> 4][tid & 15] = 0;
		ts.run[tid >> 4][tid & 15] = 0;
	}
	if (tid == 0) ts.err = 0;
	__syncthreads();
	for (int s = tid; s < nlit; s += nthr) {
		const int L = sh.lens[s];
		if (L) atomicAdd(&ts.cnt[0][L], 1u);
	}
	for (int s = tid; s < ndist; s += nthr) {
		const int L = sh.lens[nlit + s];
		if (L) atomicAdd(&ts.cnt[1][L], 1u);
	}
	__syncthreads();
	if (tid < 2) {
		const int k = tid, R = k ? kDistRoot : kLitRoot;
		Canon &cn = k ? sh.dist_c : sh.lit_c;
		int left = 1;
		for (int L = 1; L <= 15; L++) {
			left = (left << 1) - (int)ts.cnt[k][L];
			if (left < 0) {
				ts.err = D_OVERSUBSCRIBED;
				break;
			}
		}
		uint32_t code = 0, off = 0;
		for (int L = 1; L <= 15; L++) {
			ts.nxt[k][L] = code;
			cn.first[L] = (uint16_t)code;
			cn.count[L] = (uint16_t)ts.cnt[k][L];
			cn.offs[L] = (uint16_t)off;
			if (L > R) off += ts.cnt[k][L];
			code = (code + ts.cnt[k][L]) << 1;
		}
	}
	__syncthreads();
	const int err = ts.err;
	if (err) return err;
	if (warp < 2) {
		// warp 0: literal/length symbols in order, warp 1: distance symbols; the canonical code of symbol s is
		// next[L] + (number of earlier symbols of the same length)
		const int k = warp, n = k ? ndist : nlit, R = k ? kDistRoot : kLitRoot;
		const uint8_t *lens = sh.lens + (k ? nlit : 0);
		uint32_t *tab = k ? sh.dist : sh.lit;
		uint16_t *sorted = k ? sh.dist_sorted : sh.lit_sorted;
		const Canon &cn = k ? sh.dist_c : sh.lit_c;
		const uint32_t size = 1u << R;
		for (int base = 0; base < n; base += 32) {
			const int s = base + lane;
			const int L = s < n ? lens[s] : 0;
			const uint32_t m = __match_any_sync(0xffffffffu, L);
			const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
			const uint32_t r0 = ts.run[k][L];
			__syncwarp();
			if (L && lane == __ffs((int)m) - 1) ts.run[k][L] = r0 + (uint32_t)__popc(m);
			__syncwarp();
			if (L) {
				const uint32_t c = ts.nxt[k][L] + r0 + rank;
				const uint32_t rev = __brev(c) >> (32 - L);
				if (L <= R) {
					const uint32_t e = k == 0 ? litlen_entry((uint32_t)s, (uint32_t)L) : dist_entry((uint32_t)s, (uint32_t)L);
					for (uint32_t i = rev; i < size; i += (1u << L)) tab[i] = e;
				} else {
					sorted[cn.offs[L] + (c - cn.first[L])] = (uint16_t)s;
					tab[rev & (size - 1u)] = mk_entry(0, K_LONG, 0, 0);
				}
			}
		}
	}
	__syncthreads();
	return 0;
}

__device__ __forceinline__ void static_lens(uint8_t *lens) { // InflaterHuffmanTree.cs:34-70
	for (int i = threadIdx.x; i < 320; i += blockDim.x) lens[i] = (uint8_t)(i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : (i < 288 ? 8 : 5))));
}

// ---------------------------------------------------------------------------------------------------------
// span decode without per-lane caps.  MODE 0 counts; MODE 1 stores literals at out[o0 + o] and back-references at
// ml[nm] (output positions relative to the stream).
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ bool span_step2(const InfShared &sh, const uint32_t *words, Span &s, uint32_t limit, uint32_t end_rel,
                                           uint8_t *out, uint64_t o0, MatchTok *ml) {
	const uint32_t spos = s.pos;
	if (spos >= limit) return false;
	uint32_t v = peek32(words, spos);
	uint32_t nb;
	const uint32_t e = lane_decode_sym(v, sh.lit, kLitRoot, sh.lit_sorted, sh.lit_c, 0, nb);
	const uint32_t k = (e >> 4) & 15;
	if (k == K_LIT) {
		if (spos + nb > end_rel) { s.fl |= F_OVERRUN; return false; }
		if (MODE) out[o0 + s.o] = (uint8_t)(e >> 16);
		++s.o;
		s.pos = spos + nb;
		return true;
	}
	if (k == K_LEN) {
		const uint32_t xb = (e >> 8) & 15;
		const uint32_t len = (e >> 16) + ((v >> nb) & ((1u << xb) - 1u));
		uint32_t pos = spos + nb + xb;
		v = peek32(words, pos);
		uint32_t dnb;
		const uint32_t de = lane_decode_sym(v, sh.dist, kDistRoot, sh.dist_sorted, sh.dist_c, 1, dnb);
		const uint32_t dk = (de >> 4) & 15;
		if (dk != K_DIST) {
			if (pos + 15 > end_rel) s.fl |= F_OVERRUN; // diagnosed from bits past the end of the input
			else { s.fl |= F_ERR; s.det = dk == K_ILLEGAL ? D_REP_DIST : D_CODELEN0; }
			return false;
		}
		const uint32_t dxb = (de >> 8) & 15;
		const uint32_t dist = (de >> 16) + ((v >> dnb) & ((1u << dxb) - 1u));
		pos += dnb + dxb;
		if (pos > end_rel) { s.fl |= F_OVERRUN; return false; }
		if (MODE) {
			MatchTok t;
			t.out_pos = (uint32_t)(o0 + s.o);
			t.len = (uint16_t)len;
			t.dist = (uint16_t)(dist & 0xFFFFu); // 32768 fits
			ml[s.nm] = t;
		}
		++s.nm;
		s.o += len;
		s.pos = pos;
		return true;
	}
	if (k == K_EOB) {
		if (spos + nb > end_rel) s.fl |= F_OVERRUN;
		else { s.fl |= F_EOB; s.pos = spos + nb; }
		return false;
	}
	if (spos + 15 > end_rel) s.fl |= F_OVERRUN; // diagnosed from bits past the end of the input
	else { s.fl |= F_ERR; s.det = k == K_ILLEGAL ? D_REP_LEN : D_CODELEN0; }
	return false;
}

// ---------------------------------------------------------------------------------------------------------
// k_dec1: counted decode of segments
// ---------------------------------------------------------------------------------------------------------
struct __align__(16) Dec1Shared {
	InfShared sh;
	TabScratch ts;
	uint32_t in[kP1InSlots];
	uint32_t exitp[kP1Threads], flags[kP1Threads], dets[kP1Threads];
	uint32_t wsum_o[4], wsum_m[4];
	// header results (thread 0 -> all)
	int h_st, h_detail, h_btype, h_last, h_nlit, h_ndist;
	unsigned long long h_bitpos;
	uint32_t h_stored_len;
	uint32_t q_idx, a_base;
	int stop_lane;
};

__device__ __forceinline__ void stage_round(uint32_t *words, const uint32_t *gwords, uint32_t w0, uint32_t nwords, uint32_t nbytes) {
	for (int i = threadIdx.x; i < kP1InWords; i += blockDim.x)
		words[in_slot((uint32_t)i)] = ld_stream_word(gwords, w0 + (uint32_t)i, nwords, nbytes);
}

__global__ void __launch_bounds__(kP1Threads)
    k_dec1(const uint8_t *__restrict__ in, const int64_t *__restrict__ in_off, const int64_t *__restrict__ in_len, PSeg *__restrict__ segs,
           const uint32_t *__restrict__ seg_list, PCounters *__restrict__ ctr, const uint32_t *__restrict__ win_base,
           const uint32_t *__restrict__ cand, PRound *__restrict__ rounds, uint32_t round_cap, PBlockHdr *__restrict__ hdrs,
           uint32_t hdr_cap) {
	extern __shared__ __align__(16) uint8_t smem_raw[];
	Dec1Shared &S = *reinterpret_cast<Dec1Shared *>(smem_raw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t nseg = ctr->nseg;
	for (;;) {
		__syncthreads();
		if (tid == 0) S.q_idx = atomicAdd(&ctr->seg_next, 1u);
		__syncthreads();
		const uint32_t qi = S.q_idx;
		if (qi >= nseg) break;
		const uint32_t slot = seg_list[qi];
		PSeg &seg = segs[slot];
		const uint32_t stream = seg.stream;
		const uint32_t *gwords = reinterpret_cast<const uint32_t *>(in + in_off[stream]);
		const uint32_t nbytes = (uint32_t)in_len[stream];
		const uint32_t nwords = (nbytes + 3) >> 2;
		const uint64_t total_bits = 8ull * nbytes;
		const uint32_t wb = win_base[stream], nwin = win_base[stream + 1] - wb;

		// uniform across the CTA
		uint64_t bitpos = seg.start_bit, opos = 0, rs_bit = bitpos, rs_out = 0;
		uint32_t nm_total = 0;
		int st = B200Z_OK, detail = 0, end_kind = SEG_STOP;
		bool last = false;
		uint32_t blocks_done = 0;
		for (;;) { // blocks
			if (last) { // Inflater.cs:443-449: raw mode stops right behind the final block
				end_kind = SEG_FINAL;
				break;
			}
			if (blocks_done) { // a block boundary that is a candidate: its segment decodes on from here
				const uint64_t w = bitpos >> kFWShift;
				if (w < nwin && cand[wb + (uint32_t)w] == ((uint32_t)bitpos & kFWMask)) {
					end_kind = SEG_JOIN;
					break;
				}
			}
			rs_bit = bitpos;
			rs_out = opos;
			// ---- block header (thread 0) -----------------------------------------------------------------------
			if (tid == 0) {
				BitReader br;
				br.words = gwords;
				br.nbytes = nbytes;
				br.nwords = nwords;
				br.consumed = bitpos;
				br.widx = (uint32_t)(bitpos >> 5);
				br.bb = 0;
				br.bc = 0;
				const uint32_t sk = (uint32_t)(bitpos & 31);
				if (sk) {
					br.refill();
					br.bb >>= sk;
					br.bc -= sk;
				}
				int hst = B200Z_OK, hdet = 0, btype = 0, hlast = 0, nlit = 288, ndist = 32;
				uint32_t stored_len = 0;
				const uint32_t hdr = br.get(3);
				if (br.overrun()) {
					hst = B200Z_E_NEED_INPUT;
				} else {
					hlast = (int)(hdr & 1u);
					btype = (int)(hdr >> 1);
					InfShared &sh = S.sh;
					if (btype == 0) { // SkipToByteBoundary, LEN, NLEN (Inflater.cs:509)
						br.drop(br.bc & 7);
						const uint32_t len = br.get(16);
						const uint32_t nlen = br.get(16);
						if (br.overrun()) hst = B200Z_E_NEED_INPUT;
						else if (nlen != (len ^ 0xFFFFu)) { hst = B200Z_E_DATA; hdet = D_STORED_LEN; }
						stored_len = len;
					} else if (btype == 2) { // InflaterDynHeader.CreateStateMachine (:42-120)
						nlit = (int)br.get(5) + 257;
						ndist = (int)br.get(5) + 1;
						const int nmeta = (int)br.get(4) + 4;
						if (nlit > 286 || ndist > 30) { hst = B200Z_E_DATA; hdet = D_HDR_RANGE; }
						else {
							for (int i = 0; i < 19; i++) sh.lens[i] = 0;
							for (int i = 0; i < nmeta; i++) sh.lens[c_meta_order[i]] = (uint8_t)br.get(3);
							const int d = build_table(sh.lens, 19, 7, sh.meta, nullptr, nullptr, 2);
							if (d) { hst = B200Z_E_DATA; hdet = d; }
							const int total = nlit + ndist;
							int idx = 0;
							while (hst == B200Z_OK && idx < total) {
								br.refill();
								const uint32_t e = sh.meta[br.peek(7)];
								if (((e >> 4) & 15) == K_INVALID) { hst = B200Z_E_DATA; hdet = D_CODELEN0; break; }
								br.drop(e & 15);
								const int sym = (int)(e >> 16);
								if (sym < 16) {
									sh.lens[idx++] = (uint8_t)sym;
								} else {
									int rep;
									uint8_t v = 0;
									if (sym == 16) {
										if (idx == 0) { hst = B200Z_E_DATA; hdet = D_HDR_REPEAT0; break; }
										v = sh.lens[idx - 1];
										rep = 3 + (int)br.get(2);
									} else if (sym == 17) rep = 3 + (int)br.get(3);
									else rep = 11 + (int)br.get(7);
									if (idx + rep > total) { hst = B200Z_E_DATA; hdet = D_HDR_OVERRUN; break; }
									while (rep-- > 0) sh.lens[idx++] = v;
								}
								if (br.overrun()) { hst = B200Z_E_NEED_INPUT; break; }
							}
							if (hst == B200Z_OK && br.overrun()) hst = B200Z_E_NEED_INPUT;
							if (hst == B200Z_OK && sh.lens[256] == 0) { hst = B200Z_E_DATA; hdet = D_HDR_NO_EOB; }
						}
					} else if (btype == 3) {
						hst = B200Z_E_DATA;
						hdet = D_BLOCK_TYPE;
					}
				}
				// an error diagnosed from bits past the end of the input is "needs more input", not corrupt data
				if (hst != B200Z_OK && br.overrun()) { hst = B200Z_E_NEED_INPUT; hdet = 0; }
				S.h_st = hst;
				S.h_detail = hdet;
				S.h_btype = btype;
				S.h_last = hlast;
				S.h_nlit = nlit;
				S.h_ndist = ndist;
				S.h_bitpos = br.consumed;
				S.h_stored_len = stored_len;
			}
			__syncthreads();
			st = S.h_st;
			detail = S.h_detail;
			const int btype = S.h_btype;
			bitpos = S.h_bitpos;
			if (st != B200Z_OK) break;
			last = S.h_last != 0;
			uint32_t batch_base = 0, batch_used = kRoundBatch; // rounds of this block
			if (btype == 0) {
				// ---- stored block: recorded, copied by k_dec2 (OutputWindow.CopyStored :100-122) -----------------------
				const uint32_t stored_len = S.h_stored_len;
				const uint64_t ipos = bitpos >> 3;
				const uint64_t avail = ipos <= nbytes ? nbytes - ipos : 0;
				if (stored_len > avail) {
					st = B200Z_E_NEED_INPUT;
					break;
				}
				if (tid == 0) S.a_base = atomicAdd(&ctr->round_top, (uint32_t)kRoundBatch);
				__syncthreads();
				batch_base = S.a_base;
				if (batch_base + kRoundBatch > round_cap) {
					end_kind = SEG_POOL;
					break;
				}
				if (tid == 0) {
					PRound &r = rounds[batch_base];
					r.seg = slot;
					r.kind = RND_STORED;
					r.hdr = 0;
					r.w0 = (uint32_t)ipos;
					r.nbytes = stored_len;
					r.out_rel = opos;
					r.match_rel = nm_total;
					for (int k = 1; k < kRoundBatch; k++) rounds[batch_base + k].kind = RND_NONE;
				}
				opos += stored_len;
				bitpos = 8ull * (ipos + stored_len);
				++blocks_done;
				continue;
			}
			// ---- Huffman block: code tables ---------------------------------------------------------------------
			int nlit = S.h_nlit, ndist = S.h_ndist;
			if (btype == 1) {
				__syncthreads();
				static_lens(S.sh.lens);
				nlit = 288;
				ndist = 32;
				__syncthreads();
			}
			{
				const int d = build_tables_cta(S.sh, S.ts, nlit, ndist);
				if (d) {
					st = B200Z_E_DATA;
					detail = d;
					break;
				}
			}
			if (tid == 0) S.a_base = atomicAdd(&ctr->hdr_top, 1u);
			__syncthreads();
			const uint32_t hdr_idx = S.a_base;
			if (hdr_idx >= hdr_cap) {
				end_kind = SEG_POOL;
				break;
			}
			{
				PBlockHdr &h = hdrs[hdr_idx];
				if (tid == 0) {
					h.nlit = (uint16_t)nlit;
					h.ndist = (uint16_t)ndist;
					h.is_static = (uint16_t)(btype == 1);
				}
				for (int i = tid; i < 320; i += kP1Threads) h.lens[i] = S.sh.lens[i];
			}
			// ---- rounds -------------------------------------------------------------------------------------------
			bool in_block = true;
			while (in_block) {
				__syncthreads();
				const uint32_t w0 = (uint32_t)(bitpos >> 5);
				stage_round(S.in, gwords, w0, nwords, nbytes);
				__syncthreads();
				const uint32_t *words = S.in;
				const uint32_t r0 = (uint32_t)(bitpos & 31);
				const uint64_t remain = total_bits - ((uint64_t)w0 << 5);
				const uint32_t end_rel = remain > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)remain;
				const uint32_t limit = r0 + (uint32_t)(tid + 1) * kSubBits;
				uint32_t entry = r0 + (uint32_t)tid * kSubBits;
				uint32_t obytes = 0, nmatch = 0, flags = 0;
				bool changed = true, dead = false;
				for (int it = 0; it < kP1Threads + 2; it++) {
					Span sp;
					bool act = changed && !dead;
					if (changed) {
						sp.o = 0;
						sp.nm = 0;
						sp.fl = dead ? (uint32_t)F_DEAD : 0u;
						sp.det = 0;
						sp.pos = entry;
					}
					while (__any_sync(0xffffffffu, act)) {
						if (act) act = span_step2<0>(S.sh, words, sp, limit, end_rel, nullptr, 0, nullptr);
						__syncwarp();
					}
					if (changed) {
						obytes = sp.o;
						nmatch = sp.nm;
						flags = sp.fl;
						S.exitp[tid] = sp.pos;
						S.flags[tid] = sp.fl;
						S.dets[tid] = sp.det;
					}
					__syncthreads();
					changed = false;
					if (tid > 0) {
						const uint32_t pe = S.exitp[tid - 1], pf = S.flags[tid - 1];
						const bool nd = pf != 0; // the previous lane ended the block, failed or is dead itself
						changed = (pe != entry) || (nd != dead);
						entry = pe;
						dead = nd;
					}
					if (!__syncthreads_or(changed ? 1 : 0)) break;
				}
				// lanes up to and including the first one that stopped the block are exact; the rest are dead
				if (tid == 0) S.stop_lane = kP1Threads - 1;
				__syncthreads();
				if (flags & (F_EOB | F_ERR | F_OVERRUN)) atomicMin(&S.stop_lane, tid);
				__syncthreads();
				const int lastlane = S.stop_lane;
				if (tid > lastlane) {
					obytes = 0;
					nmatch = 0;
				}
				// exclusive prefix sums over the 128 lanes
				uint32_t io = obytes, im = nmatch;
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t t1 = __shfl_up_sync(0xffffffffu, io, o);
					const uint32_t t2 = __shfl_up_sync(0xffffffffu, im, o);
					if (lane >= o) {
						io += t1;
						im += t2;
					}
				}
				if (lane == 31) {
					S.wsum_o[warp] = io;
					S.wsum_m[warp] = im;
				}
				__syncthreads();
				uint32_t bo = 0, bm = 0, tot = 0, totm = 0;
				for (int w = 0; w < kP1Threads / 32; w++) {
					if (w < warp) {
						bo += S.wsum_o[w];
						bm += S.wsum_m[w];
					}
					tot += S.wsum_o[w];
					totm += S.wsum_m[w];
				}
				const uint32_t lflags = S.flags[lastlane], ldet = S.dets[lastlane], lexit = S.exitp[lastlane];
				if (batch_used == kRoundBatch) {
					if (tid == 0) S.a_base = atomicAdd(&ctr->round_top, (uint32_t)kRoundBatch);
					__syncthreads();
					batch_base = S.a_base;
					batch_used = 0;
					if (batch_base + kRoundBatch > round_cap) {
						end_kind = SEG_POOL;
						break;
					}
					if (tid < kRoundBatch) rounds[batch_base + tid].kind = RND_NONE;
					__syncthreads();
				}
				{
					PRound &r = rounds[batch_base + batch_used];
					r.entry[tid] = entry;
					r.opre[tid] = bo + io - obytes;
					r.mpre[tid] = bm + im - nmatch;
					if (tid == 0) {
						r.seg = slot;
						r.hdr = hdr_idx;
						r.w0 = w0;
						r.r0 = r0;
						r.end_rel = end_rel;
						r.lastlane = (uint32_t)lastlane;
						r.nbytes = tot;
						r.out_rel = opos;
						r.match_rel = nm_total;
						r.kind = RND_HUFF;
					}
					++batch_used;
				}
				opos += tot;
				nm_total += totm;
				bitpos = ((uint64_t)w0 << 5) + lexit;
				if (lflags & F_EOB) in_block = false;
				else if (lflags & F_OVERRUN) { st = B200Z_E_NEED_INPUT; in_block = false; }
				else if (lflags & F_ERR) { st = B200Z_E_DATA; detail = (int)ldet; in_block = false; }
				else if (tot == 0 && lexit == r0) { st = B200Z_E_INTERNAL; in_block = false; } // cannot happen: a symbol always fits
			}
			if (st != B200Z_OK || end_kind == SEG_POOL) break;
			++blocks_done;
		}
		if (tid == 0) {
			seg.end_bit = bitpos;
			seg.rs_bit = rs_bit;
			seg.rs_out = rs_out;
			seg.out_bytes = opos;
			seg.n_match = nm_total;
			seg.end_kind = end_kind;
			seg.status = st | (detail << 8);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// k_chain: one thread per stream
// ---------------------------------------------------------------------------------------------------------
__global__ void k_chain(int n, PSeg *__restrict__ segs, const uint32_t *__restrict__ win_base, const int64_t *__restrict__ in_len,
                        const int64_t *__restrict__ out_cap, const uint32_t *__restrict__ match_cap, const int32_t *__restrict__ pre,
                        int64_t *__restrict__ out_len, int64_t *__restrict__ in_used, int32_t *__restrict__ status,
                        int64_t *__restrict__ restart, uint32_t *__restrict__ str_nm, int32_t *__restrict__ fallback) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	fallback[i] = 0;
	str_nm[i] = 0;
	if (pre && pre[i] != B200Z_OK) { // the framing header was rejected (k_wrap_head): nothing to decode
		status[i] = pre[i];
		out_len[i] = 0;
		if (in_used) in_used[i] = 0;
		restart[2 * i] = 0;
		restart[2 * i + 1] = 0;
		return;
	}
	const uint32_t wb = win_base[i];
	uint64_t out = 0, nm = 0;
	uint32_t cur = (uint32_t)i;
	int st = B200Z_OK;
	uint64_t end_bit = 0, rs_bit = 0, rs_out = 0;
	bool pool = false;
	// pass 1: totals
	for (;;) {
		const PSeg &s = segs[cur];
		rs_bit = s.rs_bit;
		rs_out = out + s.rs_out;
		out += s.out_bytes;
		nm += s.n_match;
		end_bit = s.end_bit;
		if (s.end_kind == SEG_JOIN) {
			cur = (uint32_t)n + wb + (uint32_t)(s.end_bit >> kFWShift);
			continue;
		}
		if (s.end_kind == SEG_POOL) pool = true;
		else if (s.end_kind == SEG_FINAL) st = B200Z_OK;
		else st = s.status;
		break;
	}
	if (pool || out > 0xFFFF0000ull || nm > (uint64_t)match_cap[i]) {
		fallback[i] = 1; // the serial kernel decodes this stream
		return;
	}
	if (out > (uint64_t)out_cap[i]) {
		status[i] = B200Z_E_NOMEM;
		out_len[i] = 0;
		if (in_used) in_used[i] = 0;
		restart[2 * i] = (int64_t)segs[i].start_bit;
		restart[2 * i + 1] = 0;
		return;
	}
	// pass 2: positions
	cur = (uint32_t)i;
	uint64_t o = 0;
	uint32_t m = 0;
	for (;;) {
		PSeg &s = segs[cur];
		s.out_base = o;
		s.match_base = m;
		s.valid = 1;
		o += s.out_bytes;
		m += s.n_match;
		if (s.end_kind != SEG_JOIN) break;
		cur = (uint32_t)n + wb + (uint32_t)(s.end_bit >> kFWShift);
	}
	status[i] = st;
	out_len[i] = (int64_t)out;
	str_nm[i] = (uint32_t)nm;
	if (in_used) {
		uint64_t used = (end_bit + 7) >> 3; // n - RemainingInput (trap T14)
		if (used > (uint64_t)in_len[i]) used = (uint64_t)in_len[i];
		in_used[i] = (int64_t)used;
	}
	restart[2 * i] = (int64_t)rs_bit;
	restart[2 * i + 1] = (int64_t)rs_out;
}

// ---------------------------------------------------------------------------------------------------------
// k_dec2: the final decode pass, one CTA per batch of rounds
// ---------------------------------------------------------------------------------------------------------
struct __align__(16) Dec2Shared {
	InfShared sh;
	TabScratch ts;
	uint32_t in[kP1InSlots];
};

__global__ void __launch_bounds__(kP1Threads)
    k_dec2(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const int64_t *__restrict__ in_off, const int64_t *__restrict__ in_len,
           const int64_t *__restrict__ out_off, const PSeg *__restrict__ segs, const PCounters *__restrict__ ctr,
           const PRound *__restrict__ rounds, uint32_t round_cap, const PBlockHdr *__restrict__ hdrs, MatchTok *__restrict__ mlist,
           const int64_t *__restrict__ mt_off) {
	extern __shared__ __align__(16) uint8_t smem_raw[];
	Dec2Shared &S = *reinterpret_cast<Dec2Shared *>(smem_raw);
	const int tid = threadIdx.x;
	uint32_t top = ctr->round_top;
	if (top > round_cap) top = round_cap;
	for (uint32_t b0 = blockIdx.x * kRoundBatch; b0 < top; b0 += gridDim.x * kRoundBatch) {
		bool have_tab = false;
		for (int k = 0; k < kRoundBatch; k++) {
			const PRound &r = rounds[b0 + k];
			const uint32_t kind = r.kind;
			if (kind == RND_NONE) continue;
			const PSeg &seg = segs[r.seg];
			if (!seg.valid) continue;
			const uint32_t stream = seg.stream;
			uint8_t *dst = out + out_off[stream];
			const uint64_t obase = seg.out_base + r.out_rel;
			if (kind == RND_STORED) {
				const uint8_t *src = in + in_off[stream] + r.w0;
				const uint32_t len = r.nbytes;
				for (uint32_t i = tid; i < len; i += kP1Threads) dst[obase + i] = src[i];
				continue;
			}
			__syncthreads(); // the previous round's lanes are done with the staged words (and the tables)
			if (!have_tab) {
				const PBlockHdr &h = hdrs[r.hdr];
				for (int i = tid; i < 320; i += kP1Threads) S.sh.lens[i] = h.lens[i];
				__syncthreads();
				build_tables_cta(S.sh, S.ts, h.nlit, h.ndist);
				have_tab = true;
			}
			const uint32_t nbytes = (uint32_t)in_len[stream];
			stage_round(S.in, reinterpret_cast<const uint32_t *>(in + in_off[stream]), r.w0, (nbytes + 3) >> 2, nbytes);
			__syncthreads();
			{
				const bool mine = (uint32_t)tid <= r.lastlane;
				Span sp;
				sp.o = 0;
				sp.nm = 0;
				sp.fl = 0;
				sp.det = 0;
				sp.pos = mine ? r.entry[tid] : 0u;
				const uint32_t lim = r.r0 + (uint32_t)(tid + 1) * kSubBits;
				const uint64_t o0 = obase + (mine ? r.opre[tid] : 0u);
				MatchTok *ml = mlist + mt_off[stream] + seg.match_base + r.match_rel + (mine ? r.mpre[tid] : 0u);
				const uint32_t end_rel = r.end_rel;
				// warp-synchronous: the lanes step together and re-converge every symbol
				bool act = mine;
				while (__any_sync(0xffffffffu, act)) {
					if (act) act = span_step2<1>(S.sh, S.in, sp, lim, end_rel, dst, o0, ml);
					__syncwarp();
				}
			}
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------------------
// k_resolve: back-references, one CTA per stream, 16 KiB tiles in stream order
// ---------------------------------------------------------------------------------------------------------
struct __align__(16) ResShared {
	uint8_t val[kResTile];
	uint16_t ptr[kResTile];
	MatchTok mch[kResThreads];
};

__global__ void __launch_bounds__(kResThreads)
    k_resolve(const uint8_t *__restrict__ in, uint8_t *out, const int64_t *__restrict__ in_off, const int64_t *__restrict__ out_off,
              const uint32_t *__restrict__ dict_len, const int64_t *__restrict__ out_len, const uint32_t *__restrict__ str_nm,
              const int32_t *__restrict__ fallback, const MatchTok *__restrict__ mlist, const int64_t *__restrict__ mt_off, int n) {
	extern __shared__ __align__(16) uint8_t smem_raw[];
	ResShared &S = *reinterpret_cast<ResShared *>(smem_raw);
	const int stream = blockIdx.x;
	if (stream >= n || fallback[stream]) return;
	const uint32_t nm = str_nm[stream];
	if (nm == 0) return;
	const int tid = threadIdx.x;
	uint8_t *dst = out + out_off[stream];
	const int64_t total = out_len[stream];
	const MatchTok *ml = mlist + mt_off[stream];
	const uint32_t D = dict_len[stream];
	const uint8_t *dict = in + in_off[stream] - D; // the preset dictionary's tail lies in front of the compressed bytes
	uint32_t cur_m = 0;
	for (int64_t T0 = 0; T0 < total && cur_m < nm; T0 += kResTile) {
		const int tl = (int)(total - T0 < (int64_t)kResTile ? total - T0 : (int64_t)kResTile);
		const int64_t T1 = T0 + tl;
		if ((int64_t)ml[cur_m].out_pos >= T1) continue; // no back-reference touches this tile: the literals are in place
		__syncthreads();
		// the tile as k_dec2 left it (literals in place), every byte its own source
		if (tid * 16 + 16 <= tl) {
			*reinterpret_cast<uint4 *>(S.val + tid * 16) = *reinterpret_cast<const uint4 *>(dst + T0 + tid * 16);
		} else {
			for (int i = tid * 16; i < tl && i < tid * 16 + 16; i++) S.val[i] = dst[T0 + i];
		}
		for (int i = tid; i < kResTile; i += kResThreads) S.ptr[i] = (uint16_t)i;
		__syncthreads();
		// back-references that reach into the tile, 1024 at a time, each expanded by a group of eight lanes
		for (;;) {
			const uint32_t mi = cur_m + (uint32_t)tid;
			MatchTok m;
			m.out_pos = 0xFFFFFFFFu;
			m.len = 0;
			m.dist = 0;
			if (mi < nm) m = ml[mi];
			const bool in_tile = mi < nm && (int64_t)m.out_pos < T1;
			const bool done = in_tile && (int64_t)m.out_pos + m.len <= T1;
			const int n_in = __syncthreads_count(in_tile ? 1 : 0);
			const int n_done = __syncthreads_count(done ? 1 : 0);
			S.mch[tid] = m;
			__syncthreads();
			const int grp = tid >> 3, sub = tid & 7;
			for (int j = grp; j < n_in; j += kResThreads / 8) {
				const MatchTok g = S.mch[j];
				const int len = g.len;
				const int dist = g.dist ? (int)g.dist : 65536; // (never 0 for a decoded back-reference; guards the modulo)
				const int64_t dl64 = (int64_t)g.out_pos - T0;
				const int dl = (int)dl64; // may be negative: the reference started in the previous tile
				int k = sub;
				if (dl < 0) k += (-dl) & ~7; // first step whose bytes can lie in the tile
				for (; k < len; k += 8) {
					const int p = dl + k;
					if (p < 0) continue;
					if (p >= tl) break;
					const int so = dist >= len ? k : k % dist; // OutputWindow.Repeat: byte k comes from source byte k mod distance
					const int sl = dl - dist + so;
					if (sl >= 0) {
						S.ptr[p] = S.ptr[sl]; // (an ancestor of sl, whatever the other groups have written so far)
					} else {
						const int64_t ab = T0 + sl;
						uint8_t b = 0;
						if (ab >= 0) b = dst[ab];                      // final: an earlier tile
						else if (-ab <= (int64_t)D) b = dict[(int64_t)D + ab]; // preset dictionary (OutputWindow.CopyDict)
						S.val[p] = b;                                  // (else: a fresh window holds zeros, trap T13)
					}
				}
			}
			cur_m += (uint32_t)n_done;
			__syncthreads();
			if (!(n_in == kResThreads && n_done == kResThreads)) break;
		}
		// pointer jumping until every byte points at a byte that is its own source
		for (;;) {
			int changed = 0;
			for (int i = tid; i < tl; i += kResThreads) {
				const uint16_t q = S.ptr[i];
				if (q != (uint16_t)i) {
					const uint16_t r = S.ptr[q];
					if (r != q) {
						S.ptr[i] = r;
						changed = 1;
					}
				}
			}
			if (!__syncthreads_or(changed)) break;
		}
		// write the tile back
		if (tid * 16 + 16 <= tl) {
			uint32_t w[4];
			for (int q = 0; q < 4; q++) {
				uint32_t x = 0;
				for (int b = 0; b < 4; b++) x |= (uint32_t)S.val[S.ptr[tid * 16 + q * 4 + b]] << (8 * b);
				w[q] = x;
			}
			*reinterpret_cast<uint4 *>(dst + T0 + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
		} else {
			for (int i = tid * 16; i < tl && i < tid * 16 + 16; i++) dst[T0 + i] = S.val[S.ptr[i]];
		}
	}
}

} // namespace b200z
