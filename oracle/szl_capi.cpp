// szl_capi.cpp -- ORACLE (test infrastructure): flat C API over the restated classes so that
// tests/ and bench.py's cpu_baseline leg can drive them through ctypes.  The product never links this.
//
// The drain loops mirror the reference's own callers:
//   Streams/DeflaterOutputStream.cs:100-139 (Finish), :245-275 (Deflate), :388-393 (Flush), :506-510 (Write)
//   Streams/InflaterInputStream.cs:658-690 (Read)
#include "szl_oracle.hpp"
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

using namespace szl;

static thread_local std::string g_err;

#define SZL_TRY try {
#define SZL_CATCH                                                                                                      \
	}                                                                                                                  \
	catch (const SzlError &e) {                                                                                        \
		g_err = e.what();                                                                                              \
		return e.kind;                                                                                                 \
	}                                                                                                                  \
	catch (const std::exception &e) {                                                                                  \
		g_err = e.what();                                                                                              \
		return E_INTERNAL;                                                                                             \
	}                                                                                                                  \
	return 0;

struct DeflaterH {
	Deflater d;
	std::vector<uint8_t> input; // the reference keeps a reference to the caller's array; we keep a copy alive
	DeflaterH(int level, bool nowrap) : d(level, nowrap) {}
};
struct InflaterH {
	Inflater i;
	std::vector<uint8_t> input;
	explicit InflaterH(bool nowrap) : i(nowrap) {}
};

extern "C" {

const char *szl_last_error() { return g_err.c_str(); }

// ---- checksums ------------------------------------------------------------------------------
uint32_t szl_crc32(const uint8_t *buf, uint64_t n) {
	Crc32 c;
	c.Update(buf, 0, (size_t)n);
	return c.Value();
}
// running forms: `value` is the checksum's Value over the bytes so far (Crc32: 0 for none, Adler32: 1)
uint32_t szl_crc32_update(uint32_t value, const uint8_t *buf, uint64_t n) {
	Crc32 c;
	c.HarnessSetValue(value);
	c.Update(buf, 0, (size_t)n);
	return c.Value();
}
uint32_t szl_adler32(const uint8_t *buf, uint64_t n) {
	Adler32 a;
	a.Update(buf, 0, (size_t)n);
	return a.Value();
}
uint32_t szl_adler32_update(uint32_t value, const uint8_t *buf, uint64_t n) {
	Adler32 a;
	a.HarnessSetValue(value);
	a.Update(buf, 0, (size_t)n);
	return a.Value();
}
// byte-at-a-time forms (IChecksum.Update(int))
uint32_t szl_crc32_bytewise(const uint8_t *buf, uint64_t n) {
	Crc32 c;
	for (uint64_t i = 0; i < n; i++) c.Update((int)buf[i]);
	return c.Value();
}
uint32_t szl_adler32_bytewise(const uint8_t *buf, uint64_t n) {
	Adler32 a;
	for (uint64_t i = 0; i < n; i++) a.Update((int)buf[i]);
	return a.Value();
}

// ---- the test inputs of the reference's own tests ---------------------------------------------
// .NET's seeded System.Random (Knuth's subtractive generator; `new Random(seed).NextBytes(buf)`), which the reference's
// tests use for their data: Utils.GetDummyBytes (test/.../TestSupport/Utils.cs:79-85, seed 5) and the 256 MiB Adler-32
// known-answer test (test/.../Checksum/ChecksumTests.cs:41-64, seed 1).  Not part of SharpZipLib: restated from the
// published algorithm (Numerical Recipes ran3 as .NET seeds it) and pinned by that Adler-32 value in tests/test_oracle.py.
void szl_dotnet_random_bytes(int32_t seed, uint8_t *out, uint64_t n) {
	const int32_t MBIG = 2147483647, MSEED = 161803398;
	int32_t sa[56] = {0};
	const int32_t sub = seed == INT32_MIN ? MBIG : (seed < 0 ? -seed : seed);
	int32_t mj = MSEED - sub, mk = 1;
	sa[55] = mj;
	for (int i = 1; i < 55; i++) {
		const int ii = (21 * i) % 55;
		sa[ii] = mk;
		mk = mj - mk;
		if (mk < 0) mk += MBIG;
		mj = sa[ii];
	}
	for (int k = 1; k < 5; k++)
		for (int i = 1; i < 56; i++) {
			// the subtraction wraps like C#'s unchecked int arithmetic
			sa[i] = (int32_t)((uint32_t)sa[i] - (uint32_t)sa[1 + (i + 30) % 55]);
			if (sa[i] < 0) sa[i] += MBIG;
		}
	int inext = 0, inextp = 21;
	for (uint64_t j = 0; j < n; j++) {
		if (++inext >= 56) inext = 1;
		if (++inextp >= 56) inextp = 1;
		int32_t r = (int32_t)((uint32_t)sa[inext] - (uint32_t)sa[inextp]);
		if (r == MBIG) r--;
		if (r < 0) r += MBIG;
		sa[inext] = r;
		out[j] = (uint8_t)r; // NextBytes: (byte)InternalSample()
	}
}

// ---- Deflater handle ------------------------------------------------------------------------
int szl_deflater_new(int level, int nowrap, void **out) {
	SZL_TRY
	*out = new DeflaterH(level, nowrap != 0);
	SZL_CATCH
}
void szl_deflater_free(void *h) { delete (DeflaterH *)h; }
int szl_deflater_reset(void *h) {
	SZL_TRY((DeflaterH *)h)->d.Reset();
	SZL_CATCH
}
int szl_deflater_set_input(void *h, const uint8_t *buf, int32_t off, int32_t count) {
	SZL_TRY
	DeflaterH *d = (DeflaterH *)h;
	if (!d->d.IsNeedingInput()) throw SzlError(E_STATE, "Old input was not completely processed");
	if (buf == nullptr) throw SzlError(E_ARG, "buffer");
	if (off < 0 || count < 0) throw SzlError(E_ARG, "offset/count");
	d->input.assign(buf + off, buf + off + count);
	static const uint8_t empty = 0;
	d->d.SetInput(d->input.empty() ? &empty : d->input.data(), 0, count);
	SZL_CATCH
}
int szl_deflater_set_level(void *h, int level) {
	SZL_TRY((DeflaterH *)h)->d.SetLevel(level);
	SZL_CATCH
}
int szl_deflater_get_level(void *h) { return ((DeflaterH *)h)->d.GetLevel(); }
int szl_deflater_set_strategy(void *h, int s) {
	SZL_TRY
	if (s < 0 || s > 2) throw SzlError(E_ARG, "strategy");
	((DeflaterH *)h)->d.SetStrategy((DeflateStrategy)s);
	SZL_CATCH
}
int szl_deflater_set_dictionary(void *h, const uint8_t *buf, int32_t off, int32_t count) {
	SZL_TRY((DeflaterH *)h)->d.SetDictionary(buf, off, count);
	SZL_CATCH
}
void szl_deflater_flush(void *h) { ((DeflaterH *)h)->d.Flush(); }
void szl_deflater_finish(void *h) { ((DeflaterH *)h)->d.Finish(); }
int szl_deflater_deflate(void *h, uint8_t *out, int32_t off, int32_t len, int32_t *produced) {
	SZL_TRY
	*produced = ((DeflaterH *)h)->d.DeflateInto(out, off, len);
	SZL_CATCH
}
int szl_deflater_needs_input(void *h) { return ((DeflaterH *)h)->d.IsNeedingInput() ? 1 : 0; }
int szl_deflater_is_finished(void *h) { return ((DeflaterH *)h)->d.IsFinished() ? 1 : 0; }
int64_t szl_deflater_total_in(void *h) { return ((DeflaterH *)h)->d.TotalIn(); }
int64_t szl_deflater_total_out(void *h) { return ((DeflaterH *)h)->d.TotalOut(); }
uint32_t szl_deflater_adler(void *h) { return (uint32_t)((DeflaterH *)h)->d.Adler(); }

// ---- Inflater handle ------------------------------------------------------------------------
int szl_inflater_new(int nowrap, void **out) {
	SZL_TRY
	*out = new InflaterH(nowrap != 0);
	SZL_CATCH
}
void szl_inflater_free(void *h) { delete (InflaterH *)h; }
int szl_inflater_reset(void *h) {
	SZL_TRY((InflaterH *)h)->i.Reset();
	SZL_CATCH
}
int szl_inflater_set_input(void *h, const uint8_t *buf, int32_t off, int32_t count) {
	SZL_TRY
	InflaterH *d = (InflaterH *)h;
	if (!d->i.IsNeedingInput()) throw SzlError(E_STATE, "Old input was not completely processed");
	if (buf == nullptr) throw SzlError(E_ARG, "buffer");
	if (off < 0 || count < 0) throw SzlError(E_ARG, "offset/count");
	d->input.assign(buf + off, buf + off + count);
	static const uint8_t empty[2] = {0, 0};
	d->i.SetInput(d->input.empty() ? empty : d->input.data(), 0, count);
	SZL_CATCH
}
int szl_inflater_set_dictionary(void *h, const uint8_t *buf, int32_t off, int32_t count) {
	SZL_TRY((InflaterH *)h)->i.SetDictionary(buf, off, count);
	SZL_CATCH
}
int szl_inflater_inflate(void *h, uint8_t *out, int32_t outlen, int32_t off, int32_t count, int32_t *produced) {
	SZL_TRY
	*produced = ((InflaterH *)h)->i.Inflate(out, outlen, off, count);
	SZL_CATCH
}
int szl_inflater_needs_input(void *h) { return ((InflaterH *)h)->i.IsNeedingInput() ? 1 : 0; }
int szl_inflater_needs_dictionary(void *h) { return ((InflaterH *)h)->i.IsNeedingDictionary() ? 1 : 0; }
int szl_inflater_is_finished(void *h) { return ((InflaterH *)h)->i.IsFinished() ? 1 : 0; }
int32_t szl_inflater_remaining_input(void *h) { return ((InflaterH *)h)->i.RemainingInput(); }
int64_t szl_inflater_total_in(void *h) { return ((InflaterH *)h)->i.TotalIn(); }
int64_t szl_inflater_total_out(void *h) { return ((InflaterH *)h)->i.TotalOut(); }
uint32_t szl_inflater_adler(void *h) { return (uint32_t)((InflaterH *)h)->i.Adler(); }

// ---- one-shot drivers (the call patterns BASELINE.md pins) -------------------------------------
// pattern 0: SetInput(chunk)* -> Finish() -> drain Deflate() with an `obuf`-byte buffer      (C1/C3/C5)
// pattern 1: DeflaterOutputStream: Write(chunk)* -> Flush() -> Finish()                      (InflaterDeflaterTests.cs:49-62)
// pattern 2: DeflaterOutputStream: Write(chunk)* -> Finish()                                 (GZip/Zip writers, C4)
// chunk == 0 means "everything in one SetInput/Write".
static int deflate_oneshot_impl(const uint8_t *in, int64_t n, int level, int nowrap, int strategy, int pattern,
                                int32_t chunk, int32_t obuf, uint8_t *out, int64_t cap, int64_t *outlen,
                                std::vector<DeflaterHuffman::BlockTrace> *trace) {
	Deflater d(level, nowrap != 0);
	d.SetStrategy((DeflateStrategy)strategy);
	d.Engine().huffman.trace = trace;
	std::vector<uint8_t> buffer_((size_t)(obuf > 0 ? obuf : 512));
	int64_t o = 0;
	auto sink = [&](int len) {
		if (o + len > cap) throw SzlError(E_ARG, "output capacity exceeded");
		std::memcpy(out + o, buffer_.data(), (size_t)len);
		o += len;
	};
	static const uint8_t empty = 0;
	int64_t step = chunk > 0 ? chunk : (n > 0x40000000 ? 0x40000000 : n);
	int64_t pos = 0;
	do {
		int64_t c = std::min<int64_t>(step, n - pos);
		d.SetInput(n ? in + pos : &empty, 0, (int)c);
		pos += c;
		if (pattern == 0) {
			// raw Deflater user: drain until more input is needed
			while (!d.IsNeedingInput()) {
				int len = d.DeflateInto(buffer_.data(), 0, (int)buffer_.size());
				if (len <= 0) break;
				sink(len);
			}
		} else {
			// DeflaterOutputStream.Write -> Deflate() (:245-275)
			while (!d.IsNeedingInput()) {
				int len = d.DeflateInto(buffer_.data(), 0, (int)buffer_.size());
				if (len <= 0) break;
				sink(len);
			}
			if (!d.IsNeedingInput()) throw SzlError(E_DATA, "DeflaterOutputStream can't deflate all input?");
		}
	} while (pos < n);
	if (pattern == 1) {
		// DeflaterOutputStream.Flush (:388-393): deflater_.Flush(); Deflate(flushing=true)
		d.Flush();
		for (;;) {
			int len = d.DeflateInto(buffer_.data(), 0, (int)buffer_.size());
			if (len <= 0) break;
			sink(len);
		}
		if (!d.IsNeedingInput()) throw SzlError(E_DATA, "DeflaterOutputStream can't deflate all input?");
	}
	// Finish (:100-139)
	d.Finish();
	while (!d.IsFinished()) {
		int len = d.DeflateInto(buffer_.data(), 0, (int)buffer_.size());
		if (len <= 0) break;
		sink(len);
	}
	if (!d.IsFinished()) throw SzlError(E_DATA, "Can't deflate all input?");
	*outlen = o;
	return 0;
}

int szl_deflate_oneshot(const uint8_t *in, int64_t n, int level, int nowrap, int strategy, int pattern, int32_t chunk,
                        int32_t obuf, uint8_t *out, int64_t cap, int64_t *outlen) {
	SZL_TRY
	deflate_oneshot_impl(in, n, level, nowrap, strategy, pattern, chunk, obuf, out, cap, outlen, nullptr);
	SZL_CATCH
}

// same, returning the per-block decisions: rows of 5 int32 {type, nsyms, storedLength, opt_len, static_len}
int szl_deflate_trace(const uint8_t *in, int64_t n, int level, int nowrap, int strategy, int pattern, int32_t chunk,
                      uint8_t *out, int64_t cap, int64_t *outlen, int32_t *rows, int32_t maxrows, int32_t *nrows) {
	SZL_TRY
	std::vector<DeflaterHuffman::BlockTrace> tr;
	deflate_oneshot_impl(in, n, level, nowrap, strategy, pattern, chunk, 512, out, cap, outlen, &tr);
	int k = 0;
	for (auto &b : tr) {
		if (k >= maxrows) break;
		rows[5 * k + 0] = b.type;
		rows[5 * k + 1] = b.nsyms;
		rows[5 * k + 2] = b.storedLength;
		rows[5 * k + 3] = b.opt_len;
		rows[5 * k + 4] = b.static_len;
		k++;
	}
	*nrows = (int32_t)tr.size();
	SZL_CATCH
}

// InflaterInputStream.Read-style whole-buffer decode (:658-690): returns status, bytes produced, unread input
// and whether the inflater finished.  `ibuf` is the size of the input staging buffer (reference: 4096).
int szl_inflate_oneshot(const uint8_t *in, int64_t n, int nowrap, int32_t ibuf, uint8_t *out, int64_t cap,
                        int64_t *outlen, int32_t *remaining, int32_t *finished) {
	*outlen = 0;
	*remaining = 0;
	*finished = 0;
	SZL_TRY
	Inflater inf(nowrap != 0);
	int64_t ipos = 0;
	int64_t o = 0;
	if (ibuf <= 0) ibuf = 4096;
	static uint8_t scratch[1];
	// InflaterInputStream.Read loop; the caller's buffer is `out`, count = cap (fed in <= 1 GiB slices)
	for (;;) {
		int want = (int)std::min<int64_t>(cap - o, 1 << 30);
		int bytesRead = want > 0 ? inf.Inflate(out + o, want, 0, want) : inf.Inflate(scratch, 1, 0, 0);
		o += bytesRead;
		*outlen = o;
		if (inf.IsFinished()) break;
		if (inf.IsNeedingDictionary()) throw SzlError(E_DATA, "Need a dictionary");
		if (inf.IsNeedingInput()) {
			if (ipos >= n) { // InflaterInputStream.Fill :486-498
				*remaining = inf.RemainingInput();
				throw SzlError(E_DATA, "Unexpected EOF");
			}
			int c = (int)std::min<int64_t>(ibuf, n - ipos);
			inf.SetInput(in + ipos, 0, c);
			ipos += c;
		} else if (bytesRead == 0) {
			if (want == 0) throw SzlError(E_ARG, "output capacity exceeded");
			throw SzlError(E_DATA, "Invalid input data"); // InflaterInputStream.cs:676
		}
	}
	*remaining = inf.RemainingInput() + (int32_t)(n - ipos);
	*finished = inf.IsFinished() ? 1 : 0;
	SZL_CATCH
}

// ---- CPU baseline: many independent buffers over T host threads (one Deflater/Inflater per thread) ----
// direction 0 = deflate (pattern 0, whole buffer per SetInput), 1 = inflate.  Layout: buffer i is
// in[in_off[i] .. in_off[i]+in_len[i]) and writes out[out_off[i] ..), out_len[i] receives the produced size.
int szl_batch(int direction, const uint8_t *in, const int64_t *in_off, const int64_t *in_len, int32_t nbuf, int level,
              int nowrap, uint8_t *out, const int64_t *out_off, const int64_t *out_cap, int64_t *out_len,
              int32_t threads) {
	if (threads < 1) threads = 1;
	std::atomic<int> next(0), status(0);
	auto worker = [&]() {
		for (;;) {
			int i = next.fetch_add(1);
			if (i >= nbuf) return;
			int rc;
			if (direction == 0) {
				rc = szl_deflate_oneshot(in + in_off[i], in_len[i], level, nowrap, 0, 0, 0, 65536, out + out_off[i],
				                         out_cap[i], &out_len[i]);
			} else {
				int32_t rem, fin;
				rc = szl_inflate_oneshot(in + in_off[i], in_len[i], nowrap, 65536, out + out_off[i], out_cap[i],
				                         &out_len[i], &rem, &fin);
			}
			if (rc != 0) status.store(rc);
		}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < threads; t++) pool.emplace_back(worker);
	worker();
	for (auto &t : pool) t.join();
	return status.load();
}

} // extern "C"
