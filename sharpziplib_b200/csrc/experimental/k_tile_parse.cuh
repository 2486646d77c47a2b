// k_tile_parse.cuh -- EXPERIMENTAL, OPT-IN: compiled into libb200z.so (#included by b200z_deflate.cu) but launched only when
// the environment holds B200Z_TILE_PARSE=1 (k_tile_parse) or =2 (k_tile_parse2) while a level 5-9 plan is built; the default
// path (k_match + k_parse_chunk) does not change with it.  Written at the end of round 1 without GPU time left: it has
// never run on a GPU.  Bit-exact against the oracle on the CUDA emulator (tests/cuda_emu) through the GPU-tier deflate tests.
// It is the round-2 starting point that DESIGN.md 7b describes, kept next to the kernels it is meant to replace.
//
// Why: tools/match_stats.cpp shows that k_match (a match table for EVERY position) walks 7.0x the chain candidates the
// reference walks at level 6 (24.6x at level 9): DeflateSlow only searches at the positions it visits (49.6 % on the bench
// mix, 18-28 % on the compressible classes), and those are the cheap ones.  This kernel lets the parse drive the search.
//
// What: one CTA per tile of kFTile positions.  The tile's window (32 KiB of history + the tile + the longest match) and the
// window's link entries are staged in shared memory exactly as k_match stages them.  Thread t owns the 32-position segment
// t of the tile; a warp is therefore one ROUND of 1024 positions (the unit k_parse_scan / k_parse_gather work on).  Every
// thread parses its segment speculatively from a clean state with parse_step() (b200z_core.cuh, unchanged), its table
// function searching ON DEMAND in the staged window with match_search() (unchanged) and remembering the answer per
// position, so that the exit -> entry hand-off passes (across the whole CTA here, through shared memory) never search a
// position twice.  The final pass emits every round's symbols into the round's slot and leaves a RoundRec per round, which
// is what k_parse_chunk leaves today; tiles are stitched by k_parse_fix with chunk = kFTile, whose table function has to fall
// back to a search in global memory where this kernel left "not computed" (mt[p].x == kFNone).
//
// Replaces: k_match + k_parse_chunk (k_links, k_parse_fix_lazy / _scan / _gather, k_plan, k_scan, k_emit stay).
// Wiring (deflate_plan_build / deflate_plan_run): tiles of kFTile instead of kTile, p->parse_chunk = kFTile, the fix-up is
// k_parse_fix_lazy (parse_round<true>: entries this kernel left "not computed" are searched in global memory).
#pragma once

constexpr int kFTile = 16384;                 // positions per CTA
constexpr int kFThreads = kFTile / kSeg;      // 512 threads, one per 32-position segment; 16 warps = 16 rounds
constexpr int kFHist = 32768;                 // history staged in front of the tile
constexpr int kFData = kFHist + kFTile + 320; // bytes of window staged (tile + max match + pad), a multiple of 16
constexpr uint32_t kFNone = 0xFFFFFFFFu;      // memo: position not searched yet
constexpr uint32_t kFNeedB = 0x80000000u;     // memo: quarter-budget result differs from the full one, search again if asked
constexpr int kFWarm = 64;                    // k_tile_parse2<1024, true>: positions parsed in front of the tile for its entry state
constexpr int kFSmem = kFData + 2 * (kFHist + kFTile) + 4 * kFTile + kFThreads * 20 + 64 + 4 * kFWarm + 32 + 64;
static_assert(kFData % 16 == 0, "staging uses 16-byte copies");
static_assert(kFSmem <= 227 * 1024, "one CTA per SM");

struct FCarry { // ParseCarry as five words in shared memory
	uint32_t p, mlen, mstart, prevAvail, last_top;
};
__device__ __forceinline__ FCarry f_pack(const ParseCarry &c) { return FCarry{c.st.p, c.st.mlen, c.st.mstart, c.st.prevAvail, c.last_top}; }
__device__ __forceinline__ ParseCarry f_unpack(const FCarry &f) {
	ParseCarry c;
	c.st.p = f.p;
	c.st.mlen = f.mlen;
	c.st.mstart = f.mstart;
	c.st.prevAvail = f.prevAvail;
	c.last_top = f.last_top;
	return c;
}

__global__ void __launch_bounds__(kFThreads, 1)
    k_tile_parse(const uint8_t *__restrict__ in, const uint16_t *__restrict__ link, uint2 *__restrict__ mt,
                 uint32_t *__restrict__ sym_local, const int64_t *__restrict__ in_off, const int64_t *__restrict__ in_len,
                 const int2 *__restrict__ tile_desc, const uint32_t *__restrict__ rnd_off, RoundRec *__restrict__ recs,
                 const uint32_t *__restrict__ hist, const int64_t *__restrict__ bias, LevelParams lp, int strategy) {
	extern __shared__ __align__(16) uint8_t smem[];
	uint8_t *s_data = smem;
	uint16_t *s_link = reinterpret_cast<uint16_t *>(smem + kFData);
	uint32_t *s_memo = reinterpret_cast<uint32_t *>(smem + kFData + 2 * (kFHist + kFTile));
	FCarry *s_exit = reinterpret_cast<FCarry *>(smem + kFData + 2 * (kFHist + kFTile) + 4 * kFTile);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int2 td = tile_desc[blockIdx.x]; // (stream, first position of the tile: a multiple of kFTile)
	const uint32_t n = (uint32_t)in_len[td.x];
	const int64_t off = in_off[td.x];
	const uint8_t *data = in + off;
	const uint16_t *lnk = link + off;
	const uint32_t t0 = (uint32_t)td.y;
	const uint32_t t1 = (n - t0 > (uint32_t)kFTile) ? t0 + kFTile : n;
	const uint32_t w0 = t0 >= (uint32_t)kFHist ? t0 - kFHist : 0u;
	uint32_t dend = t1 + 272;
	if (dend > n) dend = n;
	const uint32_t H = hist[td.x], ab = (uint32_t)bias[td.x];
	// ---- stage the window (as k_match does; w0 and the slot base are 16-byte aligned) ----
	{
		const uint32_t nbytes = dend - w0, nvec = nbytes >> 4;
		const uint4 *src = reinterpret_cast<const uint4 *>(data + w0);
		uint4 *dst = reinterpret_cast<uint4 *>(s_data);
		for (uint32_t i = tid; i < nvec; i += kFThreads) dst[i] = __ldg(src + i);
		for (uint32_t i = (nvec << 4) + tid; i < nbytes; i += kFThreads) s_data[i] = data[w0 + i];
		const uint32_t nl = t1 - w0, nlv = nl >> 3;
		const uint4 *lsrc = reinterpret_cast<const uint4 *>(lnk + w0);
		uint4 *ldst = reinterpret_cast<uint4 *>(s_link);
		for (uint32_t i = tid; i < nlv; i += kFThreads) ldst[i] = __ldg(lsrc + i);
		for (uint32_t i = (nlv << 3) + tid; i < nl; i += kFThreads) s_link[i] = lnk[w0 + i];
		for (uint32_t i = tid; i < (uint32_t)kFTile; i += kFThreads) s_memo[i] = kFNone;
	}
	__syncthreads();
	// ---- the table function: search on demand, once per position ----
	auto tabf = [&](uint32_t p, uint32_t &a, uint32_t &b) {
		const uint32_t m = s_memo[p - t0]; // only the thread that owns p's segment ever touches this word
		if (m == kFNone) {
			match_search(s_data, s_link, w0, p, n, lp, a, b, ab);
			s_memo[p - t0] = a | (a != b ? kFNeedB : 0u);
		} else {
			a = m & ~kFNeedB;
			b = a;
			if (m & kFNeedB) { // rare: a hand-off pass asks again for a position whose quarter-budget answer differs
				LevelParams lq = lp;
				lq.chain = lp.chain >> 2;
				uint32_t dummy;
				match_search(s_data, s_link, w0, p, n, lq, b, dummy, ab);
			}
		}
	};
	auto bytef = [&](uint32_t q) { return (uint32_t)s_data[q - w0]; };
	auto slowf = [&](uint32_t p, uint32_t m0, uint32_t budget) { return match_search_above(data, lnk, p, n, m0, budget, ab); };
	// ---- speculative parse of every segment, exit -> entry hand-off across the CTA until nothing changes ----
	const uint32_t seg0 = t0 + (uint32_t)tid * kSeg, seg_end = seg0 + kSeg;
	const uint32_t lim = seg_end < n ? seg_end : n;
	// thread 0's entry is exact for the first tile of a stream (Reset state, also right behind a dictionary or a flush) and a
	// guess otherwise; k_parse_fix repairs the guess
	ParseCarry entry = clean_carry(seg0 > H ? seg0 : H), ex = entry;
	uint32_t cnt = 0;
	bool changed = true;
	for (int it = 0; it < kFThreads + 2; it++) {
		if (changed) {
			ex = entry;
			cnt = 0;
		}
		bool act = changed && ex.st.p < lim;
		while (__any_sync(0xffffffffu, act)) { // the lanes of a warp step together (a per-lane loop would leave them diverged)
			if (act) {
				ex.last_top = ex.st.p;
				uint32_t s2;
				cnt += (uint32_t)parse_step(ex.st, n, lp, strategy, tabf, bytef, slowf, s2);
				act = ex.st.p < lim;
			}
			__syncwarp();
		}
		s_exit[tid] = f_pack(ex);
		__syncthreads();
		changed = false;
		if (tid > 0) {
			const ParseCarry ne = f_unpack(s_exit[tid - 1]);
			changed = !carry_equal(ne, entry);
			entry = ne;
		}
		if (!__syncthreads_or(changed ? 1 : 0)) break; // (also keeps s_exit stable until everybody has read it)
	}
	// ---- final pass: every warp (= round) emits its symbols at prefix-summed offsets into the round's slot ----
	const uint32_t rbase = t0 + (uint32_t)warp * kRound;
	if (rbase < n) { // warp-uniform
		uint32_t incl = cnt;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		uint32_t idx = incl - cnt;
		uint32_t *sround = sym_local + off + rbase;
		ParseCarry c = entry;
		bool act = c.st.p < lim;
		while (__any_sync(0xffffffffu, act)) {
			if (act) {
				c.last_top = c.st.p;
				uint32_t s2;
				if (parse_step(c.st, n, lp, strategy, tabf, bytef, slowf, s2)) sround[idx++] = s2;
				act = c.st.p < lim;
			}
			__syncwarp();
		}
		const ParseCarry last = shfl_carry(c, 31);
		const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
		if (lane == 0) {
			RoundRec r;
			r.p = last.st.p;
			r.mlen = last.st.mlen;
			r.mstart = last.st.mstart;
			r.prevAvail = last.st.prevAvail;
			r.last_top = last.last_top;
			r.cnt = total;
			(recs + rnd_off[td.x])[rbase / kRound] = r;
		}
	}
	__syncthreads();
	// ---- what was searched goes to the table for the fix-up; the rest is marked "not computed" ----
	uint2 *out = mt + off + t0;
	for (uint32_t i = tid; i < t1 - t0; i += kFThreads) {
		const uint32_t m = s_memo[i];
		out[i] = m == kFNone ? make_uint2(kFNone, kFNone) : make_uint2(m & ~kFNeedB, (m & kFNeedB) ? kFNone : (m & ~kFNeedB));
	}
}

// ------------------------------------------------------------------------------------------------------------------
// k_tile_parse2 -- "guess, batch, re-parse".  k_tile_parse above lets every lane search for itself, and the lock-step model
// (tools/match_stats.cpp) predicts 18 % busy lanes: a warp iteration costs what its slowest lane walks.  Here a segment
// that needs an entry nobody has computed yet takes a 1-hop PROXY for it (one candidate, cheap), notes the position in the
// tile's request list and parses on; at the end of the pass the whole CTA serves the list as one batch (every thread takes
// requests, all walks start together -- the shape k_match has), then the segments that used a proxy or whose entry changed
// parse again.  Same tool, same workload: 5.6 passes per tile on average (11 at most), batches of 45 % / 3.7 % / 1.0 % /
// 0.2 % ... of the tile's positions, 11.9 candidates per position in balanced batches instead of k_match's 34.3, parse
// work of 0.08 lock-step iterations per position: about 2.4x less issue work than k_match + k_parse together.
// Interface as k_tile_parse plus `scratch` (4 bytes per position that are free until k_parse_gather: the request list).
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kFReq = 0xFFFFFFFEu; // memo: requested, the batch has not run yet

// k_match's chain walk (b200z_deflate.cu: bytes 2..9 of the scan string in registers, 4-byte extension compares on the
// staged window, one budget test per candidate), as a function over a window whose byte 0 is stream position w0 and whose
// link entries start at w0 as well.  Same results as match_search(); the byte-wise loop of match_search() ran with ~2 of 32
// lanes active in k_match and took ~30 % of it.
__device__ __forceinline__ void tile_walk(const uint8_t *s_data, const uint16_t *s_link, uint32_t w0, uint32_t p, uint32_t n,
                                          const LevelParams &lp, uint32_t ab, uint32_t &resA, uint32_t &resB) {
	resA = 0;
	resB = 0;
	const uint32_t la = n - p;
	if (la < (uint32_t)kMinMatch) return;
	uint32_t d = (uint32_t)s_link[p - w0];
	if (d > (uint32_t)kMaxDist - (is_slide_pos(p + ab) ? 1u : 0u)) d = 0; // DeflaterEngine.cs:788 + trap T8
	if (d == 0) return;
	const uint32_t chain = (uint32_t)lp.chain, budgetB = chain >> 2;
	const uint32_t maxlen = la < (uint32_t)kMaxMatch ? la : (uint32_t)kMaxMatch;
	const uint32_t nice = la < (uint32_t)lp.nice ? la : (uint32_t)lp.nice;
	const uint32_t is = p - w0;
	const uint8_t *sp = s_data + is;
	uint32_t m = kMinMatch - 1, bd = 0, dist = d, cnt = 0;
	bool haveB = false;
	uint32_t stop = budgetB ? budgetB : chain;
	const uint32_t s0 = sp[0], s1 = sp[1];
	uint32_t scan_end1 = s1, scan_end = sp[2];
	uint32_t sw0, sw1;
	{
		const uint32_t as = is + 2;
		const uint32_t *ws = reinterpret_cast<const uint32_t *>(s_data + (as & ~3u));
		const uint32_t w1 = ws[1], sh = (as & 3u) * 8u;
		sw0 = __funnelshift_r(ws[0], w1, sh);
		sw1 = __funnelshift_r(w1, ws[2], sh);
	}
	for (;;) {
		const uint32_t ic = is - dist;
		const uint8_t *c = s_data + ic;
		++cnt;
		if (c[m] == scan_end && c[m - 1] == scan_end1 && c[0] == s0 && c[1] == s1) {
			uint32_t l = 2;
			if (maxlen >= 10) {
				const uint32_t ac = ic + 2;
				const uint32_t *wc = reinterpret_cast<const uint32_t *>(s_data + (ac & ~3u));
				const uint32_t w1 = wc[1], sh = (ac & 3u) * 8u;
				uint32_t x = __funnelshift_r(wc[0], w1, sh) ^ sw0;
				if (x) {
					l = 2 + ((uint32_t)(__ffs((int)x) - 1) >> 3);
					goto lcp_done;
				}
				x = __funnelshift_r(w1, wc[2], sh) ^ sw1;
				if (x) {
					l = 6 + ((uint32_t)(__ffs((int)x) - 1) >> 3);
					goto lcp_done;
				}
				l = 10;
			}
			while (l + 4 <= maxlen) {
				const uint32_t ac = ic + l, as = is + l;
				const uint32_t *wc = reinterpret_cast<const uint32_t *>(s_data + (ac & ~3u));
				const uint32_t *ws = reinterpret_cast<const uint32_t *>(s_data + (as & ~3u));
				const uint32_t x = __funnelshift_r(wc[0], wc[1], (ac & 3u) * 8u) ^ __funnelshift_r(ws[0], ws[1], (as & 3u) * 8u);
				if (x) {
					l += (uint32_t)(__ffs((int)x) - 1) >> 3;
					goto lcp_done;
				}
				l += 4;
			}
			while (l < maxlen && c[l] == sp[l]) ++l;
		lcp_done:
			if (l > m) {
				m = l;
				bd = dist;
				if (m >= nice) break;
				scan_end1 = sp[m - 1];
				scan_end = sp[m];
			}
		}
		if (cnt == stop) { // one test per candidate: first the quarter budget (B's snapshot), then the full one
			if (stop == chain) break;
			resB = m >= (uint32_t)kMinMatch ? pack_match(m, bd) : 0u;
			haveB = true;
			stop = chain;
		}
		const uint32_t l2 = s_link[is - dist];
		if (l2 == 0) break;
		dist += l2;
		if (dist >= (uint32_t)kMaxDist) break; // chain entries need distance < 32506 (T7)
	}
	resA = m >= (uint32_t)kMinMatch ? pack_match(m, bd) : 0u;
	if (!haveB) resB = resA;
}

// kThreads = kFThreads (512): every thread owns a segment.  kThreads = 1024: the upper 16 warps own no segment and wait at the
// barrier while the others parse; staging, the batch phase (classification, ordering, walks) and the table write-back run on
// all 32 warps -- the walks are the bulk of the work and k_match needed its 32 warps per SM to hide the shared-memory hops
// (64 registers at 1024 threads, no spills).
// kWarm (1024 threads only): the tile's entry state is not the clean guess but what a parse of the kFWarm positions in front
// of the tile ends in -- warp 16, idle while the others parse, searches those positions (two per lane, exact) and its lane 0
// parses them; thread 0 takes the result like a hand-off from a segment -1.  tools/tile_fixup.cpp: the clean guess is wrong
// at 92 % of the tile boundaries (one round parsed again by the fix-up each time), a 64-position warm-up leaves 0.2 %.  The
// entry the tile finally used goes to `ent` (one record per round, only tile starts written) so that the fix-up compares the
// true state with it instead of with the clean guess.
template <int kThreads, bool kWarm>
__global__ void __launch_bounds__(kThreads, 1)
    k_tile_parse2(const uint8_t *__restrict__ in, const uint16_t *__restrict__ link, uint2 *__restrict__ mt,
                  uint32_t *__restrict__ sym_local, const int64_t *__restrict__ in_off, const int64_t *__restrict__ in_len,
                  const int2 *__restrict__ tile_desc, const uint32_t *__restrict__ rnd_off, RoundRec *__restrict__ recs,
                  const uint32_t *__restrict__ hist, const int64_t *__restrict__ bias, uint32_t *__restrict__ scratch, LevelParams lp,
                  int strategy, RoundRec *__restrict__ ent) {
	static_assert(!kWarm || kThreads == 1024, "the warm-up runs on a warp that owns no segment");
	extern __shared__ __align__(16) uint8_t smem[];
	uint8_t *s_data = smem;
	uint16_t *s_link = reinterpret_cast<uint16_t *>(smem + kFData);
	uint32_t *s_memo = reinterpret_cast<uint32_t *>(smem + kFData + 2 * (kFHist + kFTile));
	FCarry *s_exit = reinterpret_cast<FCarry *>(smem + kFData + 2 * (kFHist + kFTile) + 4 * kFTile);
	__shared__ uint32_t s_nreq;
	uint32_t *s_warm = reinterpret_cast<uint32_t *>(smem + kFSmem - 64 - 4 * kFWarm - 32); // kWarm: exact entries of the warm-up positions
	FCarry *s_wexit = reinterpret_cast<FCarry *>(smem + kFSmem - 64 - 32);                // kWarm: where the warm-up parse ended
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int2 td = tile_desc[blockIdx.x];
	const uint32_t n = (uint32_t)in_len[td.x];
	const int64_t off = in_off[td.x];
	const uint8_t *data = in + off;
	const uint16_t *lnk = link + off;
	const uint32_t t0 = (uint32_t)td.y;
	const uint32_t t1 = (n - t0 > (uint32_t)kFTile) ? t0 + kFTile : n;
	const uint32_t w0 = t0 >= (uint32_t)kFHist ? t0 - kFHist : 0u;
	uint32_t dend = t1 + 272;
	if (dend > n) dend = n;
	const uint32_t H = hist[td.x], ab = (uint32_t)bias[td.x];
	// the tile's 4 x kFTile bytes of scratch: request list (2 bytes per position: relative to t0, at most one entry per
	// position), the same list ordered by walk-length class, and the class of every request
	uint16_t *req = reinterpret_cast<uint16_t *>(scratch + off + t0);
	uint16_t *ord = req + (t1 - t0); // (a last, partial tile only owns 4 x (t1 - t0) bytes of scratch)
	__shared__ uint32_t s_cls[2][kMatchClasses];
	uint8_t *req_cls = reinterpret_cast<uint8_t *>(mt + off + t0); // the tile's table slots are written at the very end: 8 bytes per position free until then
	{
		const uint32_t nbytes = dend - w0, nvec = nbytes >> 4;
		const uint4 *src = reinterpret_cast<const uint4 *>(data + w0);
		uint4 *dst = reinterpret_cast<uint4 *>(s_data);
		for (uint32_t i = tid; i < nvec; i += kThreads) dst[i] = __ldg(src + i);
		for (uint32_t i = (nvec << 4) + tid; i < nbytes; i += kThreads) s_data[i] = data[w0 + i];
		const uint32_t nl = t1 - w0, nlv = nl >> 3;
		const uint4 *lsrc = reinterpret_cast<const uint4 *>(lnk + w0);
		uint4 *ldst = reinterpret_cast<uint4 *>(s_link);
		for (uint32_t i = tid; i < nlv; i += kThreads) ldst[i] = __ldg(lsrc + i);
		for (uint32_t i = (nlv << 3) + tid; i < nl; i += kThreads) s_link[i] = lnk[w0 + i];
		for (uint32_t i = tid; i < (uint32_t)kFTile; i += kThreads) s_memo[i] = kFNone;
		if (tid == 0) s_nreq = 0;
	}
	__syncthreads();
	LevelParams l1 = lp; // the proxy: the first chain candidate only
	l1.chain = 1;
	bool used_proxy = false;
	auto tabf = [&](uint32_t p, uint32_t &a, uint32_t &b) {
		const uint32_t m = s_memo[p - t0]; // only the owner of p's segment reads or writes this word outside the batch phase
		if (m < kFReq) {
			a = m & ~kFNeedB;
			b = a;
			if (m & kFNeedB) {
				LevelParams lq = lp;
				lq.chain = lp.chain >> 2;
				uint32_t dummy;
				match_search(s_data, s_link, w0, p, n, lq, b, dummy, ab);
			}
			return;
		}
		if (m == kFNone) {
			if (s_link[p - w0] == 0) { // no chain: the exact answer is "nothing", no request needed
				s_memo[p - t0] = 0;
				a = b = 0;
				return;
			}
			s_memo[p - t0] = kFReq;
			req[atomicAdd(&s_nreq, 1u)] = (uint16_t)(p - t0);
		}
		used_proxy = true;
		match_search(s_data, s_link, w0, p, n, l1, a, b, ab);
	};
	auto bytef = [&](uint32_t q) { return (uint32_t)s_data[q - w0]; };
	auto slowf = [&](uint32_t p, uint32_t m0, uint32_t budget) { return match_search_above(data, lnk, p, n, m0, budget, ab); };
	const bool owner = kThreads == kFThreads || tid < kFThreads; // (warp-uniform) this thread owns a segment
	const uint32_t seg0 = t0 + (uint32_t)(owner ? tid : 0) * kSeg, seg_end = seg0 + kSeg;
	const uint32_t lim = !owner ? 0u : seg_end < n ? seg_end : n;
	ParseCarry entry = clean_carry(seg0 > H ? seg0 : H), ex = entry;
	uint32_t cnt = 0;
	bool dirty = owner;
	// warm-up range [ws, t0): behind the history (the state at H is exactly clean) and only for tiles that start on a guess
	const uint32_t ws = (!kWarm || t0 == 0 || t0 <= H) ? t0 : (t0 - H > (uint32_t)kFWarm ? t0 - kFWarm : H);
	const bool has_warm = kWarm && ws < t0;
	if (kWarm && has_warm && warp == kFThreads / 32) { // (warp-uniform) the first warp without segments
		for (uint32_t p = ws + (uint32_t)lane; p < t0; p += 32) {
			uint32_t a, b;
			tile_walk(s_data, s_link, w0, p, n, lp, ab, a, b);
			s_warm[p - ws] = a | (a != b ? kFNeedB : 0u);
		}
		__syncwarp();
		if (lane == 0) {
			auto tabw = [&](uint32_t p, uint32_t &a, uint32_t &b) {
				const uint32_t m = s_warm[p - ws];
				a = m & ~kFNeedB;
				b = a;
				if (m & kFNeedB) {
					LevelParams lq = lp;
					lq.chain = lp.chain >> 2;
					uint32_t dummy;
					match_search(s_data, s_link, w0, p, n, lq, b, dummy, ab);
				}
			};
			ParseCarry c = clean_carry(ws);
			while (c.st.p < t0) {
				c.last_top = c.st.p;
				uint32_t s2;
				parse_step(c.st, n, lp, strategy, tabw, bytef, slowf, s2);
			}
			*s_wexit = f_pack(c);
		}
		__syncwarp();
	}
	for (int pass = 0; pass < 4 * kFThreads; pass++) {
		// ---- parse: exact entries where there are, proxies elsewhere ----
		if (dirty) {
			ex = entry;
			cnt = 0;
			used_proxy = false;
		}
		bool act = dirty && ex.st.p < lim;
		while (__any_sync(0xffffffffu, act)) {
			if (act) {
				ex.last_top = ex.st.p;
				uint32_t s2;
				cnt += (uint32_t)parse_step(ex.st, n, lp, strategy, tabf, bytef, slowf, s2);
				act = ex.st.p < lim;
			}
			__syncwarp();
		}
		if (owner) s_exit[tid] = f_pack(ex);
		__syncthreads();
		// ---- batch: the CTA serves the pass's requests ----
		const uint32_t nreq = s_nreq;
		const uint16_t *work = req;
		if (nreq >= 4u * kFThreads) {
			// a big batch (the first one holds ~45 % of the tile) is ordered by expected walk length, longest first, the way
			// k_match orders its tile (first eight hops exact, then their density extrapolated over the window): the lanes of
			// a warp then hold walks of similar length.  The order only decides who computes what.
			if (tid < 2 * kMatchClasses) (&s_cls[0][0])[tid] = 0;
			__syncthreads();
			for (uint32_t r = tid; r < nreq; r += kThreads) {
				const uint32_t is = (uint32_t)req[r] + (t0 - w0);
				uint32_t dist = s_link[is], hops = 1, cls;
				while (hops < 8) {
					const uint32_t l2 = s_link[is - dist];
					if (l2 == 0 || dist + l2 >= (uint32_t)kMaxDist) break;
					dist += l2;
					++hops;
				}
				if (hops < 8) cls = hops <= 1 ? 1u : hops <= 2 ? 2u : hops <= 4 ? 3u : hops <= 6 ? 4u : 5u;
				else {
					uint32_t est = (8u * (uint32_t)kMaxDist) / dist;
					if (est > (uint32_t)lp.chain) est = (uint32_t)lp.chain;
					cls = est <= 11 ? 6u : est <= 16 ? 7u : est <= 23 ? 8u : est <= 32 ? 9u : est <= 45 ? 10u : est <= 64 ? 11u
					      : est <= 91 ? 12u : est <= 128 ? 13u : est <= 512 ? 14u : 15u;
				}
				req_cls[r] = (uint8_t)cls;
				atomicAdd(&s_cls[0][cls], 1u);
			}
			__syncthreads();
			if (tid == 0) {
				uint32_t b = 0;
				for (int c = kMatchClasses - 1; c >= 1; c--) {
					s_cls[1][c] = b;
					b += s_cls[0][c];
				}
			}
			__syncthreads();
			for (uint32_t r = tid; r < nreq; r += kThreads) ord[atomicAdd(&s_cls[1][req_cls[r]], 1u)] = req[r];
			__syncthreads();
			work = ord;
		}
		for (uint32_t r = tid; r < nreq; r += kThreads) {
			const uint32_t i = work[r];
			uint32_t a, b;
			tile_walk(s_data, s_link, w0, t0 + i, n, lp, ab, a, b);
			s_memo[i] = a | (a != b ? kFNeedB : 0u);
		}
		__syncthreads();
		if (tid == 0) s_nreq = 0;
		// ---- who parses again: a proxy was used (its entry is exact now), or the predecessor's exit moved ----
		bool changed = false;
		if (owner && tid > 0) {
			const ParseCarry ne = f_unpack(s_exit[tid - 1]);
			changed = !carry_equal(ne, entry);
			entry = ne;
		} else if (kWarm && has_warm && tid == 0) { // (written before the first barrier of pass 0, never again)
			const ParseCarry ne = f_unpack(*s_wexit);
			changed = !carry_equal(ne, entry);
			entry = ne;
		}
		dirty = changed || used_proxy;
		if (!__syncthreads_or(dirty ? 1 : 0)) break;
	}
	// ---- final pass (every entry it reads is exact) and the table for the fix-up: as k_tile_parse ----
	if (kWarm && tid == 0) { // the state this tile's records and symbols start from
		RoundRec r;
		r.p = entry.st.p;
		r.mlen = entry.st.mlen;
		r.mstart = entry.st.mstart;
		r.prevAvail = entry.st.prevAvail;
		r.last_top = entry.last_top;
		r.cnt = has_warm ? 1u : 0u; // 0: the clean guess (compared without last_top, as the fix-up always did)
		(ent + rnd_off[td.x])[t0 / kRound] = r;
	}
	const uint32_t rbase = t0 + (uint32_t)warp * kRound;
	if (owner && rbase < n) {
		uint32_t incl = cnt;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		uint32_t idx = incl - cnt;
		uint32_t *sround = sym_local + off + rbase;
		ParseCarry c = entry;
		bool act = c.st.p < lim;
		while (__any_sync(0xffffffffu, act)) {
			if (act) {
				c.last_top = c.st.p;
				uint32_t s2;
				if (parse_step(c.st, n, lp, strategy, tabf, bytef, slowf, s2)) sround[idx++] = s2;
				act = c.st.p < lim;
			}
			__syncwarp();
		}
		const ParseCarry last = shfl_carry(c, 31);
		const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
		if (lane == 0) {
			RoundRec r;
			r.p = last.st.p;
			r.mlen = last.st.mlen;
			r.mstart = last.st.mstart;
			r.prevAvail = last.st.prevAvail;
			r.last_top = last.last_top;
			r.cnt = total;
			(recs + rnd_off[td.x])[rbase / kRound] = r;
		}
	}
	__syncthreads();
	uint2 *out = mt + off + t0;
	for (uint32_t i = tid; i < t1 - t0; i += kThreads) {
		const uint32_t m = s_memo[i];
		out[i] = m >= kFReq ? make_uint2(kFNone, kFNone) : make_uint2(m & ~kFNeedB, (m & kFNeedB) ? kFNone : (m & ~kFNeedB));
	}
}
