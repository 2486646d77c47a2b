// b200z_api.cu -- the C-ABI of libb200z.so (include/b200z.h): context, plans, host-buffer batch calls and the
// streaming handles that mirror Deflater.cs / Inflater.cs member for member.  No CPU codec lives here: every byte of
// compressed or decompressed data is produced by the kernels in b200z_deflate.cu / b200z_inflate.cu.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "b200z_internal.cuh"

namespace b200z {

static thread_local std::string g_err;
static std::mutex g_mu;
static int g_device = -1;

void set_error(const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_err = buf;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
	set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
	return B200Z_E_CUDA;
}

int ensure_init() {
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_device >= 0) return B200Z_OK;
	int cnt = 0;
	cudaError_t e = cudaGetDeviceCount(&cnt);
	if (e != cudaSuccess || cnt == 0) {
		set_error("no CUDA device: libb200z has no CPU fallback (%s)", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
		return B200Z_E_CUDA;
	}
	int dev = 0;
	B200Z_CUDA(cudaGetDevice(&dev));
	g_device = dev;
	return B200Z_OK;
}

int Arena::alloc() {
	size = align_up(used + 256, 256);
	cudaError_t e = cudaMalloc(&base, (size_t)size);
	if (e != cudaSuccess) {
		base = nullptr;
		set_error("cudaMalloc of %lld workspace bytes failed: %s", (long long)size, cudaGetErrorString(e));
		return B200Z_E_NOMEM;
	}
	return B200Z_OK;
}
void Arena::release() {
	if (base) cudaFree(base);
	base = nullptr;
}

// ---- static tables blob ---------------------------------------------------------------------------
static void fill_static_blob(uint8_t *b) {
	// encoder side: DeflaterHuffman static ctor (:602-642); decoder side: InflaterHuffmanTree static ctor (:34-70)
	size_t o = 0;
	for (int i = 0; i < kLiteralNum; i++) {
		uint16_t c = (uint16_t)static_lcode(i);
		memcpy(b + o, &c, 2);
		o += 2;
	}
	for (int i = 0; i < kLiteralNum; i++) b[o++] = (uint8_t)static_llen(i);
	for (int i = 0; i < kDistNum; i++) {
		uint16_t c = (uint16_t)static_dcode(i);
		memcpy(b + o, &c, 2);
		o += 2;
	}
	for (int i = 0; i < kDistNum; i++) b[o++] = 5;
	for (int i = 0; i < 288; i++) b[o++] = (uint8_t)(i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8)));
	for (int i = 0; i < 32; i++) b[o++] = 5;
}
constexpr int kStaticBlob = kLiteralNum * 3 + kDistNum * 3 + 288 + 32;

// ---- small host helpers ----------------------------------------------------------------------------
struct HostBits { // carries the sub-byte tail of a flushed stream on the host (PendingBuffer's `bits`, :23-24)
	uint32_t bits = 0;
	int count = 0;
	void put(std::vector<uint8_t> &out, uint32_t v, int n) {
		bits |= v << count;
		count += n;
		while (count >= 8) {
			out.push_back((uint8_t)bits);
			bits >>= 8;
			count -= 8;
		}
	}
	void align(std::vector<uint8_t> &out) {
		if (count > 0) out.push_back((uint8_t)bits);
		bits = 0;
		count = 0;
	}
};

struct PinnedBuf {
	uint8_t *p = nullptr;
	size_t cap = 0;
	int ensure(size_t n) {
		if (n <= cap) return B200Z_OK;
		if (p) cudaFreeHost(p);
		p = nullptr;
		cap = 0;
		cudaError_t e = cudaHostAlloc((void **)&p, n, cudaHostAllocDefault);
		if (e != cudaSuccess) {
			set_error("cudaHostAlloc of %zu bytes failed: %s", n, cudaGetErrorString(e));
			return B200Z_E_NOMEM;
		}
		cap = n;
		return B200Z_OK;
	}
	~PinnedBuf() {
		if (p) cudaFreeHost(p);
	}
};

struct DevBuf {
	uint8_t *p = nullptr;
	size_t cap = 0;
	int ensure(size_t n) {
		if (n <= cap) return B200Z_OK;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		cudaError_t e = cudaMalloc((void **)&p, n);
		if (e != cudaSuccess) {
			set_error("cudaMalloc of %zu bytes failed: %s", n, cudaGetErrorString(e));
			return B200Z_E_NOMEM;
		}
		cap = n;
		return B200Z_OK;
	}
	~DevBuf() {
		if (p) cudaFree(p);
	}
};

// Runs one plan end to end from host buffers: pinned staging, H2D, kernels, D2H.
struct HostRunResult {
	std::vector<int64_t> out_len, in_used;
	std::vector<int32_t> status;
	std::vector<uint32_t> check;
};

static int run_plan_host(b200z_plan *plan, const uint8_t *const *in, PinnedBuf &hin, PinnedBuf &hout, DevBuf &din,
                         DevBuf &dout, DevBuf &dmeta, HostRunResult &r, bool fetch_all_out, const uint32_t *check_seed = nullptr) {
	const int n = plan->n;
	int rc;
	if ((rc = hin.ensure((size_t)plan->in_bytes + 256))) return rc;
	if ((rc = din.ensure((size_t)plan->in_bytes + 256))) return rc;
	if ((rc = dout.ensure((size_t)plan->out_bytes + 256))) return rc;
	const size_t meta_bytes = (size_t)n * (8 + 8 + 4 + 4) + 256;
	if ((rc = dmeta.ensure(meta_bytes))) return rc;
	for (int i = 0; i < n; i++)
		if (plan->in_len[i]) memcpy(hin.p + plan->in_off[i], in[i], (size_t)plan->in_len[i]);
	cudaStream_t s = 0;
	B200Z_CUDA(cudaMemcpyAsync(din.p, hin.p, (size_t)plan->in_bytes, cudaMemcpyHostToDevice, s));
	int64_t *d_out_len = reinterpret_cast<int64_t *>(dmeta.p);
	int64_t *d_in_used = d_out_len + n;
	int32_t *d_status = reinterpret_cast<int32_t *>(d_in_used + n);
	uint32_t *d_check = reinterpret_cast<uint32_t *>(d_status + n);
	if (check_seed) B200Z_CUDA(cudaMemcpyAsync(d_check, check_seed, 4ull * n, cudaMemcpyHostToDevice, s)); // running values
	rc = b200z_plan_run(plan, din.p, dout.p, d_out_len, d_status, d_check, d_in_used, (void *)s);
	if (rc) return rc;
	r.out_len.assign(n, 0);
	r.in_used.assign(n, 0);
	r.status.assign(n, 0);
	r.check.assign(n, 0);
	B200Z_CUDA(cudaMemcpyAsync(r.out_len.data(), d_out_len, 8ull * n, cudaMemcpyDeviceToHost, s));
	B200Z_CUDA(cudaMemcpyAsync(r.in_used.data(), d_in_used, 8ull * n, cudaMemcpyDeviceToHost, s));
	B200Z_CUDA(cudaMemcpyAsync(r.status.data(), d_status, 4ull * n, cudaMemcpyDeviceToHost, s));
	B200Z_CUDA(cudaMemcpyAsync(r.check.data(), d_check, 4ull * n, cudaMemcpyDeviceToHost, s));
	if ((rc = hout.ensure((size_t)plan->out_bytes + 256))) return rc;
	if (fetch_all_out) {
		B200Z_CUDA(cudaMemcpyAsync(hout.p, dout.p, (size_t)plan->out_bytes, cudaMemcpyDeviceToHost, s));
		B200Z_CUDA(cudaStreamSynchronize(s));
	} else {
		B200Z_CUDA(cudaStreamSynchronize(s));
		for (int i = 0; i < n; i++) {
			if (r.out_len[i] > 0)
				B200Z_CUDA(cudaMemcpyAsync(hout.p + plan->out_off[i], dout.p + plan->out_off[i], (size_t)r.out_len[i],
				                           cudaMemcpyDeviceToHost, s));
		}
		B200Z_CUDA(cudaStreamSynchronize(s));
	}
	return B200Z_OK;
}

} // namespace b200z

using namespace b200z;

extern "C" {

const char *b200z_last_error(void) { return g_err.c_str(); }
int b200z_version(void) { return 100; }

int b200z_init(int device) {
	{
		std::lock_guard<std::mutex> lk(g_mu);
		int cnt = 0;
		cudaError_t e = cudaGetDeviceCount(&cnt);
		if (e != cudaSuccess || cnt == 0) {
			set_error("no CUDA device: libb200z has no CPU fallback (%s)", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
			return B200Z_E_CUDA;
		}
		if (device < 0 || device >= cnt) {
			set_error("device %d out of range (%d devices)", device, cnt);
			return B200Z_E_ARG;
		}
		B200Z_CUDA(cudaSetDevice(device));
		g_device = device;
	}
	return checksum_init_tables();
}

int b200z_static_tables_size(void) { return kStaticBlob; }
int b200z_static_tables_export(uint8_t *blob, int32_t cap) {
	if (!blob || cap < kStaticBlob) {
		set_error("static table blob needs %d bytes", kStaticBlob);
		return B200Z_E_ARG;
	}
	fill_static_blob(blob);
	return B200Z_OK;
}
int b200z_static_tables_import(const uint8_t *blob, int32_t len) {
	// The kernels derive the static codes arithmetically (b200z_core.cuh static_lcode/static_llen), so installing a
	// broadcast copy reduces to verifying that the sender's tables are the ones this rank would use.
	if (!blob || len != kStaticBlob) {
		set_error("static table blob has %d bytes, expected %d", len, kStaticBlob);
		return B200Z_E_ARG;
	}
	uint8_t mine[kStaticBlob];
	fill_static_blob(mine);
	if (memcmp(mine, blob, kStaticBlob) != 0) {
		set_error("static Huffman tables received from the root differ from the local ones");
		return B200Z_E_DATA;
	}
	return B200Z_OK;
}

int64_t b200z_deflate_bound(int64_t len) { return len + (len >> 3) + 1024; }
int64_t b200z_engine_state_bytes(void) { return kFastStateBytes; }

// ---- plans -------------------------------------------------------------------------------------------
int b200z_deflate_plan_create(int32_t n, const int64_t *in_len, int level, int strategy, int wrap, int end_mode,
                              b200z_plan **plan) {
	return b200z_deflate_plan_create_ex(n, in_len, level, strategy, wrap, end_mode, nullptr, plan);
}

int b200z_deflate_plan_create_ex(int32_t n, const int64_t *in_len, int level, int strategy, int wrap, int end_mode,
                                 const b200z_history *hist, b200z_plan **plan) {
	if (!plan || n < 0 || (n > 0 && !in_len)) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	if (level == -1) level = 6;
	if (level < 0 || level > 9) {
		set_error("level");
		return B200Z_E_ARG;
	}
	if (strategy < 0 || strategy > 2 || wrap < 0 || wrap > B200Z_WRAP_RAW_CRC32 || end_mode < 0 || end_mode > 2) {
		set_error("strategy/wrap/end_mode");
		return B200Z_E_ARG;
	}
	int rc = ensure_init();
	if (rc) return rc;
	b200z_plan *p = new b200z_plan();
	p->kind = 0;
	p->n = n;
	p->level = level;
	p->strategy = strategy;
	p->wrap = wrap;
	p->end_mode = end_mode;
	p->in_len.assign(in_len, in_len + n);
	if (hist && n > 0 && level <= 4) {
		// levels 0-4: the call pattern and the engine state between segments (header: "streams with history")
		if (hist->chunk_count) {
			if (!hist->chunk_len) {
				set_error("history: chunk_count without chunk_len");
				delete p;
				return B200Z_E_ARG;
			}
			p->sched_cum.resize(n);
			for (int i = 0; i < n; i++) {
				int64_t sum = 0;
				for (int k = 0; k < hist->chunk_count[i]; k++) {
					const int64_t c = hist->chunk_len[i] ? hist->chunk_len[i][k] : -1;
					if (c < 0) {
						set_error("stream %d: SetInput size %d", i, k);
						delete p;
						return B200Z_E_ARG;
					}
					sum += c;
					p->sched_cum[i].push_back((uint32_t)sum);
				}
				if (hist->chunk_count[i] < 0 || (hist->chunk_count[i] > 0 && sum != in_len[i])) {
					set_error("stream %d: the SetInput sizes add up to %lld, the stream has %lld bytes", i, (long long)sum,
					          (long long)in_len[i]);
					delete p;
					return B200Z_E_ARG;
				}
			}
		}
		if (hist->undrained_last) p->undrained.assign(hist->undrained_last, hist->undrained_last + n);
		if (hist->engine_state && level >= 1) p->engine_state.assign(hist->engine_state, hist->engine_state + n);
		if (hist->stored_state && level == 0) p->stored_state = hist->stored_state;
	}
	if (hist && hist->kind != B200Z_HIST_NONE && n > 0) {
		if ((hist->kind != B200Z_HIST_DICTIONARY && hist->kind != B200Z_HIST_CONTINUE) || !hist->hist_len) {
			set_error("history: kind/hist_len");
			delete p;
			return B200Z_E_ARG;
		}
		if (hist->kind == B200Z_HIST_CONTINUE && level < 5) {
			// DeflateStored / DeflateFast keep window-relative state across Deflate() calls that is not a function of the
			// stream position: it has to come from the run of the previous segment
			bool have = level == 0 ? p->stored_state != nullptr : !p->engine_state.empty();
			for (void *q : p->engine_state) have = have && q != nullptr;
			if (!have) {
				set_error("continuing a stream after Flush() at levels 0-4 needs the engine state of the previous segment "
				          "(b200z_history.engine_state / stored_state)");
				delete p;
				return B200Z_E_UNSUPPORTED;
			}
		}
		p->hist_kind = hist->kind;
		p->check_seeded = hist->check_seeded != 0;
		p->hist.assign(hist->hist_len, hist->hist_len + n);
		p->pos_base.resize(n);
		p->bit_base.resize(n);
		p->hist_mask.resize(n);
		for (int i = 0; i < n; i++) {
			const int64_t H = p->hist[i];
			if (H < 0 || H > 32768 || (hist->kind == B200Z_HIST_DICTIONARY && H > kMaxDist)) {
				set_error("stream %d: history of %lld bytes (at most %d)", i, (long long)H,
				          hist->kind == B200Z_HIST_DICTIONARY ? kMaxDist : 32768);
				delete p;
				return B200Z_E_ARG;
			}
			p->pos_base[i] = (hist->pos_base && hist->kind == B200Z_HIST_CONTINUE) ? hist->pos_base[i] : H;
			p->bit_base[i] = hist->bit_base ? hist->bit_base[i] : 0;
			if (p->pos_base[i] < H || p->pos_base[i] > 0xFFFF0000ll || p->bit_base[i] < 0 || p->bit_base[i] > 7) {
				set_error("stream %d: pos_base/bit_base", i);
				delete p;
				return B200Z_E_ARG;
			}
			if (hist->hist_mask && hist->hist_mask[i]) p->hist_mask[i].assign(hist->hist_mask[i], hist->hist_mask[i] + H);
			p->in_len[i] += H; // the input slot holds history + data
		}
	}
	rc = deflate_plan_build(p);
	if (rc) {
		p->ws.release();
		delete p;
		return rc;
	}
	*plan = p;
	return B200Z_OK;
}

int b200z_inflate_plan_create(int32_t n, const int64_t *comp_len, const int64_t *out_cap, int wrap, b200z_plan **plan) {
	return b200z_inflate_plan_create_ex(n, comp_len, out_cap, wrap, nullptr, plan);
}

int b200z_inflate_plan_create_ex(int32_t n, const int64_t *comp_len, const int64_t *out_cap, int wrap, const int64_t *dict_len,
                                 b200z_plan **plan) {
	if (!plan || n < 0 || (n > 0 && (!comp_len || !out_cap)) || wrap < 0 || wrap > B200Z_WRAP_RAW_CRC32) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	int rc = ensure_init();
	if (rc) return rc;
	b200z_plan *p = new b200z_plan();
	p->kind = 1;
	p->n = n;
	p->wrap = wrap;
	p->in_len.assign(comp_len, comp_len + n);
	p->out_cap.assign(out_cap, out_cap + n);
	if (dict_len && n > 0) {
		p->hist_kind = B200Z_HIST_DICTIONARY;
		p->hist.assign(dict_len, dict_len + n);
		for (int i = 0; i < n; i++) {
			if (p->hist[i] < 0 || p->hist[i] > 32768) { // OutputWindow.CopyDict keeps the last WindowSize bytes (:160-166)
				set_error("stream %d: dictionary of %lld bytes (at most 32768)", i, (long long)p->hist[i]);
				delete p;
				return B200Z_E_ARG;
			}
			p->in_len[i] += p->hist[i];
		}
	}
	rc = inflate_plan_build(p);
	if (rc) {
		p->ws.release();
		delete p;
		return rc;
	}
	*plan = p;
	return B200Z_OK;
}

int b200z_inflate_plan_set_start_bits(b200z_plan *plan, const int32_t *start_bit) {
	if (!plan || plan->kind != 1 || (plan->n > 0 && !start_bit)) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	if (plan->wrap != B200Z_WRAP_RAW && plan->wrap != B200Z_WRAP_RAW_CRC32) {
		set_error("start bits are for raw streams (a framed stream starts behind its header)");
		return B200Z_E_ARG;
	}
	std::vector<uint32_t> sb((size_t)plan->n);
	for (int i = 0; i < plan->n; i++) {
		if (start_bit[i] < 0 || start_bit[i] > 7) {
			set_error("stream %d: start bit %d (0..7)", i, start_bit[i]);
			return B200Z_E_ARG;
		}
		sb[(size_t)i] = (uint32_t)start_bit[i];
	}
	if (plan->n) B200Z_CUDA(cudaMemcpy(plan->ws.at<uint32_t>(plan->o_start_bit), sb.data(), 4ull * plan->n, cudaMemcpyHostToDevice));
	plan->has_start_bits = true;
	return B200Z_OK;
}

int b200z_plan_get_restart_points(b200z_plan *plan, int64_t *bit, int64_t *out_pos, void *cuda_stream) {
	if (!plan || plan->kind != 1 || (plan->n > 0 && (!bit || !out_pos))) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	if (plan->n == 0) return B200Z_OK;
	std::vector<int64_t> rp(2 * (size_t)plan->n);
	cudaStream_t s = (cudaStream_t)cuda_stream;
	B200Z_CUDA(cudaMemcpyAsync(rp.data(), plan->ws.at<int64_t>(plan->o_restart), 16ull * plan->n, cudaMemcpyDeviceToHost, s));
	B200Z_CUDA(cudaStreamSynchronize(s));
	for (int i = 0; i < plan->n; i++) {
		bit[i] = rp[2 * (size_t)i];
		out_pos[i] = rp[2 * (size_t)i + 1];
	}
	return B200Z_OK;
}

int b200z_plan_set_timing(b200z_plan *plan, int enable) {
	if (!plan) return B200Z_E_ARG;
	plan->timing = enable != 0;
	plan->ev_used = 0;
	return B200Z_OK;
}

int b200z_plan_get_timings(b200z_plan *plan, char *names, int32_t names_cap, float *ms, int32_t cap, int32_t *count) {
	if (!plan || !ms || !count) return B200Z_E_ARG;
	*count = 0;
	std::string nm;
	for (int i = 0; i + 1 < plan->ev_used && *count < cap; i++) {
		float t = 0;
		cudaError_t e = cudaEventElapsedTime(&t, plan->ev[i], plan->ev[i + 1]);
		if (e != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime", __FILE__, __LINE__);
		ms[*count] = t;
		nm += plan->ev_name[i];
		nm += ";";
		++*count;
	}
	if (names && names_cap > 0) {
		strncpy(names, nm.c_str(), (size_t)names_cap - 1);
		names[names_cap - 1] = 0;
	}
	return B200Z_OK;
}

int b200z_plan_destroy(b200z_plan *plan) {
	if (!plan) return B200Z_OK;
	for (cudaEvent_t e : plan->ev) cudaEventDestroy(e);
	plan->ws.release();
	delete plan;
	return B200Z_OK;
}
int64_t b200z_plan_in_bytes(const b200z_plan *p) { return p->in_bytes; }
int64_t b200z_plan_out_bytes(const b200z_plan *p) { return p->out_bytes; }
int64_t b200z_plan_in_offset(const b200z_plan *p, int32_t i) { return p->in_off[i]; }
int64_t b200z_plan_data_offset(const b200z_plan *p, int32_t i) {
	return p->in_off[i] + (p->hist.empty() ? 0 : p->hist[i]); // behind the history / dictionary
}
int64_t b200z_plan_out_offset(const b200z_plan *p, int32_t i) { return p->out_off[i]; }
int64_t b200z_plan_out_capacity(const b200z_plan *p, int32_t i) { return p->out_cap[i]; }
int64_t b200z_plan_workspace_bytes(const b200z_plan *p) { return p->ws.size; }
int32_t b200z_plan_launches(const b200z_plan *p) { return p->launches; }

int b200z_plan_run(b200z_plan *plan, const uint8_t *d_in, uint8_t *d_out, int64_t *d_out_len, int32_t *d_status,
                   uint32_t *d_check, int64_t *d_in_used, void *cuda_stream) {
	return b200z_plan_run_stages(plan, d_in, d_out, d_out_len, d_status, d_check, d_in_used, B200Z_STAGE_SEARCH | B200Z_STAGE_ENCODE,
	                             cuda_stream);
}

int b200z_plan_run_stages(b200z_plan *plan, const uint8_t *d_in, uint8_t *d_out, int64_t *d_out_len, int32_t *d_status,
                          uint32_t *d_check, int64_t *d_in_used, int stages, void *cuda_stream) {
	if (!plan || !d_in || !d_out || !d_out_len || !d_status || (stages & ~3) || stages == 0) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	cudaStream_t s = (cudaStream_t)cuda_stream;
	if (plan->kind == 0) return deflate_plan_run(plan, d_in, d_out, d_out_len, d_status, d_check, d_in_used, s, stages);
	if (!(stages & B200Z_STAGE_ENCODE)) return B200Z_OK; // an inflate plan is a single stage
	return inflate_plan_run(plan, d_in, d_out, d_out_len, d_status, d_check, d_in_used, s);
}

} // extern "C" (reopened below)

// ---- packing the produced streams back to back (so a caller copies only what was produced) -----------------
namespace b200z {
__global__ void k_pack_offsets(int n, const int64_t *__restrict__ len, int64_t *__restrict__ off) {
	// n is a batch size (thousands at most): one thread block, serial carry between 1024-element chunks
	__shared__ int64_t s_part[32];
	__shared__ int64_t s_carry;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024) {
		const int i = base + threadIdx.x;
		const int64_t v = i < n ? ((len[i] + 15) & ~15ll) : 0; // 16-byte aligned starts: vector copies on both sides
		int64_t incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if ((threadIdx.x & 31) >= o) incl += t;
		}
		if ((threadIdx.x & 31) == 31) s_part[threadIdx.x >> 5] = incl;
		__syncthreads();
		int64_t woff = 0;
		for (int k = 0; k < (int)(threadIdx.x >> 5); k++) woff += s_part[k];
		const int64_t carry = s_carry;
		if (i < n) off[i] = carry + woff + incl - v;
		__syncthreads();
		if (threadIdx.x == 1023) s_carry = carry + woff + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) off[n] = s_carry;
}

__global__ void __launch_bounds__(256)
    k_pack(const uint8_t *__restrict__ out, const int64_t *__restrict__ out_off, const int64_t *__restrict__ len,
           const int64_t *__restrict__ off, uint8_t *__restrict__ packed) {
	const int i = blockIdx.x;
	const uint8_t *src = out + out_off[i];
	uint8_t *dst = packed + off[i];
	const int64_t nb = len[i];
	const int64_t nv = (nb + 15) >> 4; // the slot is 256-byte aligned and at least 16 bytes longer than the data
	const uint4 *sv = reinterpret_cast<const uint4 *>(src);
	uint4 *dv = reinterpret_cast<uint4 *>(dst);
	for (int64_t k = threadIdx.x; k < nv; k += blockDim.x) dv[k] = sv[k];
}
} // namespace b200z

extern "C" int b200z_plan_pack(b200z_plan *plan, const uint8_t *d_out, const int64_t *d_out_len, uint8_t *d_packed,
                               int64_t *d_packed_off, void *cuda_stream) {
	if (!plan || !d_out || !d_out_len || !d_packed || !d_packed_off) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	if (plan->n == 0) return B200Z_OK;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	k_pack_offsets<<<1, 1024, 0, s>>>(plan->n, d_out_len, d_packed_off);
	k_pack<<<plan->n, 256, 0, s>>>(d_out, plan->ws.at<int64_t>(plan->o_out_off), d_out_len, d_packed_off, d_packed);
	B200Z_CUDA(cudaGetLastError());
	return B200Z_OK;
}

extern "C" {

// ---- checksums ---------------------------------------------------------------------------------------
static int checksum_host(int kind, const uint8_t *buf, int64_t len, uint32_t *value) {
	if (!value || len < 0 || (len > 0 && !buf)) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	int rc = ensure_init();
	if (rc) return rc;
	if (len == 0) return B200Z_OK;
	DevBuf d, meta;
	if ((rc = d.ensure((size_t)len + 256))) return rc;
	B200Z_CUDA(cudaMemcpy(d.p, buf, (size_t)len, cudaMemcpyHostToDevice));
	std::vector<CkTile> tiles;
	checksum_tiles(&len, 1, tiles, kind);
	const size_t tb = sizeof(CkTile) * tiles.size();
	if ((rc = meta.ensure(tb + 256))) return rc;
	int64_t *d_off = reinterpret_cast<int64_t *>(meta.p);
	int64_t *d_len = d_off + 1;
	unsigned long long *d_acc = reinterpret_cast<unsigned long long *>(d_len + 1);
	uint32_t *d_val = reinterpret_cast<uint32_t *>(d_acc + 2);
	CkTile *d_tiles = reinterpret_cast<CkTile *>(meta.p + 64);
	int64_t zero = 0;
	B200Z_CUDA(cudaMemcpy(d_off, &zero, 8, cudaMemcpyHostToDevice));
	B200Z_CUDA(cudaMemcpy(d_len, &len, 8, cudaMemcpyHostToDevice));
	B200Z_CUDA(cudaMemcpy(d_val, value, 4, cudaMemcpyHostToDevice));
	B200Z_CUDA(cudaMemcpy(d_tiles, tiles.data(), tb, cudaMemcpyHostToDevice));
	rc = checksum_launch(kind, d.p, d_off, d_len, 1, d_tiles, (int32_t)tiles.size(), d_acc, d_val, 0, 0);
	if (rc) return rc;
	B200Z_CUDA(cudaMemcpy(value, d_val, 4, cudaMemcpyDeviceToHost));
	return B200Z_OK;
}
int b200z_crc32(const uint8_t *buf, int64_t len, uint32_t *value) { return checksum_host(0, buf, len, value); }
int b200z_adler32(const uint8_t *buf, int64_t len, uint32_t *value) { return checksum_host(1, buf, len, value); }

int b200z_checksum_batch_device(int kind, const uint8_t *d_data, const int64_t *off, const int64_t *len, int32_t n,
                                uint32_t *d_value, void *cuda_stream) {
	if (kind < 0 || kind > 1 || n < 0 || (n > 0 && (!d_data || !off || !len || !d_value))) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	int rc = ensure_init();
	if (rc) return rc;
	if (n == 0) return B200Z_OK;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	std::vector<CkTile> tiles;
	checksum_tiles(len, n, tiles, kind);
	const size_t tb = sizeof(CkTile) * tiles.size();
	uint8_t *meta = nullptr;
	const size_t bytes = 16ull * n + 16ull * n + tb + 256;
	B200Z_CUDA(cudaMallocAsync((void **)&meta, bytes, s));
	int64_t *d_off = reinterpret_cast<int64_t *>(meta);
	int64_t *d_len = d_off + n;
	unsigned long long *d_acc = reinterpret_cast<unsigned long long *>(d_len + n);
	CkTile *d_tiles = reinterpret_cast<CkTile *>(d_acc + 2 * n);
	B200Z_CUDA(cudaMemcpyAsync(d_off, off, 8ull * n, cudaMemcpyHostToDevice, s));
	B200Z_CUDA(cudaMemcpyAsync(d_len, len, 8ull * n, cudaMemcpyHostToDevice, s));
	if (tb) B200Z_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tb, cudaMemcpyHostToDevice, s));
	rc = checksum_launch(kind, d_data, d_off, d_len, n, d_tiles, (int32_t)tiles.size(), d_acc, d_value, 0, s);
	B200Z_CUDA(cudaStreamSynchronize(s)); // the pageable descriptor copies above must finish before `tiles` dies
	B200Z_CUDA(cudaFreeAsync(meta, s));
	return rc;
}

// ---- host-buffer batch calls ----------------------------------------------------------------------------
static void zlib_header(int level, uint8_t h[2], bool preset_dict = false) { // Deflater.cs:436-464 (trap T11)
	int header = (8 + (7 << 4)) << 8;
	int level_flags = (level - 1) >> 1;
	if (level_flags < 0 || level_flags > 3) level_flags = 3;
	header |= level_flags << 6;
	if (preset_dict) header |= 0x20; // DeflaterConstants.PRESET_DICT
	header += 31 - (header % 31);
	h[0] = (uint8_t)(header >> 8);
	h[1] = (uint8_t)header;
}

int b200z_deflate_batch(const uint8_t *const *in, const int64_t *in_len, int32_t n, int level, int strategy, int wrap,
                        int end_mode, uint8_t *const *out, const int64_t *out_cap, int64_t *out_len, uint32_t *check,
                        int32_t *status) {
	if (n < 0 || (n > 0 && (!in || !in_len || !out || !out_cap || !out_len))) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	if (wrap == B200Z_WRAP_GZIP) {
		// GZipOutputStream's header carries caller state (MTIME, FNAME; GzipOutputStream.cs:339-375): the host shim
		// writes header and trailer around the raw stream and takes the CRC32 from `check`.
		set_error("gzip framing is written by the host stream layer; deflate with wrap=B200Z_WRAP_RAW_CRC32 and use check (CRC32)");
		return B200Z_E_UNSUPPORTED;
	}
	if (level == -1) level = 6;
	b200z_plan *plan = nullptr;
	int rc = b200z_deflate_plan_create(n, in_len, level, strategy, wrap, end_mode, &plan);
	if (rc) return rc;
	PinnedBuf hin, hout;
	DevBuf din, dout, dmeta;
	HostRunResult r;
	rc = run_plan_host(plan, in, hin, hout, din, dout, dmeta, r, false);
	int first = B200Z_OK;
	if (!rc) {
		for (int i = 0; i < n; i++) {
			int st = r.status[i] & 0xFF;
			int64_t need = r.out_len[i] + (wrap == B200Z_WRAP_ZLIB ? 6 : 0);
			if (st == B200Z_OK && need > out_cap[i]) st = B200Z_E_NOMEM;
			if (st == B200Z_OK) {
				uint8_t *o = out[i];
				if (wrap == B200Z_WRAP_ZLIB) {
					zlib_header(level, o);
					o += 2;
				}
				memcpy(o, hout.p + plan->out_off[i], (size_t)r.out_len[i]);
				o += r.out_len[i];
				if (wrap == B200Z_WRAP_ZLIB) { // Adler32 trailer, big endian (Deflater.cs:509-514)
					const uint32_t a = r.check[i];
					o[0] = (uint8_t)(a >> 24);
					o[1] = (uint8_t)(a >> 16);
					o[2] = (uint8_t)(a >> 8);
					o[3] = (uint8_t)a;
				}
				out_len[i] = need;
			} else {
				out_len[i] = 0;
				if (first == B200Z_OK) first = st;
			}
			if (status) status[i] = st;
			if (check) check[i] = r.check[i];
		}
	}
	b200z_plan_destroy(plan);
	if (rc) return rc;
	if (first != B200Z_OK) set_error("stream failed with status %d", first);
	return first;
}

static const char *inflate_detail_msg(int detail);

int b200z_inflate_batch(const uint8_t *const *in, const int64_t *in_len, int32_t n, int wrap, uint8_t *const *out,
                        const int64_t *out_cap, int64_t *out_len, int64_t *in_used, uint32_t *check, int32_t *status) {
	if (n < 0 || (n > 0 && (!in || !in_len || !out || !out_cap || !out_len))) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	b200z_plan *plan = nullptr;
	int rc = b200z_inflate_plan_create(n, in_len, out_cap, wrap, &plan);
	if (rc) return rc;
	PinnedBuf hin, hout;
	DevBuf din, dout, dmeta;
	HostRunResult r;
	rc = run_plan_host(plan, in, hin, hout, din, dout, dmeta, r, false);
	int first = B200Z_OK, first_detail = 0;
	if (!rc) {
		for (int i = 0; i < n; i++) {
			const int st = r.status[i];
			if (r.out_len[i] > 0) memcpy(out[i], hout.p + plan->out_off[i], (size_t)r.out_len[i]);
			out_len[i] = r.out_len[i];
			if (in_used) in_used[i] = r.in_used[i];
			if (status) status[i] = st;
			if (check) check[i] = wrap != B200Z_WRAP_RAW ? r.check[i] : 0u;
			if ((st & 0xFF) != B200Z_OK && first == B200Z_OK) {
				first = st & 0xFF;
				first_detail = (st >> 8) & 0xFF;
			}
		}
	}
	b200z_plan_destroy(plan);
	if (rc) return rc;
	if (first == B200Z_E_DATA) set_error("%s", inflate_detail_msg(first_detail)); // the reference's exception message
	else if (first != B200Z_OK) set_error("stream failed with status %d", first);
	return first;
}

// =====================================================================================================
// Streaming handles
// =====================================================================================================
struct DeflaterH {
	int level = 6, strategy = 0;
	bool raw = false;
	// Deflater.cs state bits (:96-110)
	bool flushing = false, finishing = false, finished = false;
	bool header_done = false;
	bool flushed_once = false; // a sync flush has been emitted and nothing was compressed since
	// what the engine's window has seen so far, cut to its last 32768 bytes (preset dictionary, then every compressed
	// segment); hist_mask flags the positions InsertString never saw (the last two of the dictionary / of each segment)
	std::vector<uint8_t> history, hist_mask;
	int64_t window_seen = 0;  // dictionary bytes kept + TotalIn: the SlideWindow phase of the next segment
	bool dict_set = false, deflate_called = false, started = false;
	uint32_t dict_adler = 0;
	std::vector<uint8_t> input;    // everything SetInput handed over and not yet compressed
	std::vector<uint8_t> pending;  // produced bytes not yet drained by Deflate()
	size_t pending_pos = 0;
	HostBits tail;                 // sub-byte tail carried between device runs
	int64_t total_in = 0, total_out = 0;
	uint32_t adler = 1;
	// levels 0-4 depend on the call pattern (trap T9) and keep engine state across Flush(): the sizes of the SetInput calls
	// since the last run, whether Deflate() was called behind the last of them, and what the engines carry
	std::vector<int64_t> chunks;
	bool undrained = false;
	b200z_stored_state sstate = {0, 0, 0, 0}; // level 0 (DeflateStored)
	DevBuf fstate;                            // levels 1-4 (DeflateFast: head[], prev[], scalars), on the device
	PinnedBuf hin, hout;
	DevBuf din, dout, dmeta;
};

static void deflater_remember(DeflaterH *d, const uint8_t *seg, size_t len, size_t uninserted) {
	// append a dictionary/segment to the carried window image
	d->history.insert(d->history.end(), seg, seg + len);
	d->hist_mask.insert(d->hist_mask.end(), len, 0);
	for (size_t k = 0; k < uninserted && k < len; k++) d->hist_mask[d->hist_mask.size() - 1 - k] = 1;
	if (d->history.size() > 32768) {
		const size_t cut = d->history.size() - 32768;
		d->history.erase(d->history.begin(), d->history.begin() + (ptrdiff_t)cut);
		d->hist_mask.erase(d->hist_mask.begin(), d->hist_mask.begin() + (ptrdiff_t)cut);
	}
	d->window_seen += (int64_t)len;
}

static int deflater_run_device(DeflaterH *d, int end_mode) {
	// compresses d->input as the next segment of the stream; END_FLUSH keeps the stream open and may end inside a byte
	const int64_t len = (int64_t)d->input.size();
	const int64_t H = (int64_t)d->history.size();
	const bool continuing = d->started; // an earlier segment (possibly empty) has been emitted
	b200z_history hs;
	memset(&hs, 0, sizeof hs);
	const int32_t n_chunks = (int32_t)d->chunks.size();
	const int64_t *chunk_ptr = d->chunks.data();
	const int32_t undrained = (d->undrained && n_chunks > 0) ? 1 : 0;
	void *fstate_ptr = nullptr;
	if (d->level <= 4) {
		hs.chunk_count = &n_chunks;
		hs.chunk_len = &chunk_ptr;
		hs.undrained_last = &undrained;
		if (d->level == 0) hs.stored_state = &d->sstate;
		else {
			int rc0 = d->fstate.ensure((size_t)b200z_engine_state_bytes());
			if (rc0) return rc0;
			fstate_ptr = d->fstate.p;
			hs.engine_state = &fstate_ptr;
		}
	}
	const int64_t pos_base = d->window_seen;
	const int32_t bit_base = d->tail.count;
	const uint8_t *mask = d->hist_mask.data();
	hs.kind = continuing ? B200Z_HIST_CONTINUE : (H ? B200Z_HIST_DICTIONARY : B200Z_HIST_NONE);
	hs.check_seeded = 1;
	hs.hist_len = &H;
	hs.pos_base = &pos_base;
	hs.bit_base = &bit_base;
	hs.hist_mask = &mask;
	const uint32_t stored_before = d->sstate.input_off;
	b200z_plan *plan = nullptr;
	int rc = b200z_deflate_plan_create_ex(1, &len, d->level, d->strategy, d->raw ? B200Z_WRAP_RAW : B200Z_WRAP_ZLIB, end_mode,
	                                      (hs.kind != B200Z_HIST_NONE || d->level <= 4) ? &hs : nullptr, &plan);
	if (rc) return rc;
	std::vector<uint8_t> slot;
	const uint8_t *inp = d->input.data();
	if (H) {
		slot.reserve((size_t)(H + len));
		slot.insert(slot.end(), d->history.begin(), d->history.end());
		slot.insert(slot.end(), d->input.begin(), d->input.end());
		inp = slot.data();
	}
	HostRunResult r;
	const uint32_t seed = d->adler;
	rc = run_plan_host(plan, &inp, d->hin, d->hout, d->din, d->dout, d->dmeta, r, false,
	                   (plan->check_seeded && !d->raw) ? &seed : nullptr);
	if (!rc && (r.status[0] & 0xFF) != B200Z_OK) {
		rc = r.status[0] & 0xFF;
		set_error("device deflate failed with status %d", rc);
	}
	if (!rc) {
		const uint8_t *o = d->hout.p + plan->out_off[0];
		const int64_t bits = r.in_used[0]; // deflate plans report the exact bit length here (bit_base included)
		const int64_t whole = bits >> 3;
		std::vector<uint8_t> seg(o, o + ((bits + 7) >> 3));
		if (!seg.empty()) seg[0] |= (uint8_t)d->tail.bits; // the carried sub-byte tail completes the first byte
		d->pending.insert(d->pending.end(), seg.begin(), seg.begin() + (ptrdiff_t)whole);
		// PendingBuffer keeps the sub-byte tail in `bits` until later writes complete the byte (:168-189)
		d->tail.count = (int)(bits & 7);
		d->tail.bits = d->tail.count ? (seg[(size_t)whole] & ((1u << d->tail.count) - 1u)) : 0u;
		if (end_mode == B200Z_END_FINISH) d->tail.align(d->pending); // FINISHING_STATE: AlignToByte (:507)
		if (!d->raw) d->adler = r.check[0];
		// (level 0: Finish() behind an undrained SetInput can end the stream before all input is taken, see stored_run)
		d->total_in += d->level == 0 ? (int64_t)(uint32_t)(d->sstate.input_off - stored_before) : len;
		d->started = true;
		deflater_remember(d, d->input.data(), (size_t)len, 2);
		d->input.clear();
		d->chunks.clear();
		d->undrained = false;
	}
	b200z_plan_destroy(plan);
	return rc;
}

int b200z_deflater_create(int level, int raw, void **h) {
	if (!h) {
		set_error("h");
		return B200Z_E_ARG;
	}
	if (level == -1) level = 6;
	else if (level < 0 || level > 9) {
		set_error("level"); // ArgumentOutOfRangeException(nameof(level)), Deflater.cs:184-187
		return B200Z_E_ARG;
	}
	DeflaterH *d = new DeflaterH();
	d->level = level;
	d->raw = raw != 0;
	*h = d;
	return B200Z_OK;
}
int b200z_deflater_destroy(void *h) {
	delete (DeflaterH *)h;
	return B200Z_OK;
}
int b200z_deflater_reset(void *h) { // Deflater.Reset :204-210 keeps level and strategy
	DeflaterH *d = (DeflaterH *)h;
	d->flushing = d->finishing = d->finished = false;
	d->header_done = false;
	d->flushed_once = false;
	d->input.clear();
	d->pending.clear();
	d->pending_pos = 0;
	d->tail = HostBits();
	d->total_in = d->total_out = 0;
	d->adler = 1;
	d->history.clear();
	d->hist_mask.clear();
	d->window_seen = 0;
	d->dict_set = d->deflate_called = d->started = false;
	d->dict_adler = 0;
	d->chunks.clear();
	d->undrained = false;
	d->sstate = b200z_stored_state{0, 0, 0, 0};
	return B200Z_OK;
}
int b200z_deflater_set_level(void *h, int level) {
	DeflaterH *d = (DeflaterH *)h;
	if (level == -1) level = 6;
	else if (level < 0 || level > 9) {
		set_error("level");
		return B200Z_E_ARG;
	}
	if (level != d->level && (!d->input.empty() || d->total_in > 0)) {
		set_error("SetLevel in mid-stream (DeflaterEngine.SetLevel flushes a block, trap T17) is not accelerated");
		return B200Z_E_UNSUPPORTED;
	}
	d->level = level;
	return B200Z_OK;
}
int b200z_deflater_get_level(void *h, int *level) {
	*level = ((DeflaterH *)h)->level;
	return B200Z_OK;
}
int b200z_deflater_set_strategy(void *h, int strategy) {
	if (strategy < 0 || strategy > 2) {
		set_error("strategy");
		return B200Z_E_ARG;
	}
	((DeflaterH *)h)->strategy = strategy;
	return B200Z_OK;
}
int b200z_deflater_set_dictionary(void *h, const uint8_t *dict, int32_t len) {
	DeflaterH *d = (DeflaterH *)h;
	// Deflater.SetDictionary :372-381: only in INIT_STATE (before the header went out); the header's FDICT bit and
	// DICTID come from the engine's Adler-32 over the dictionary, which is then reset (:386-401)
	if (d->raw || d->deflate_called || d->header_done || d->total_in > 0 || d->dict_set) {
		// state != INIT_STATE (:561): a raw deflater starts in BUSY_STATE (:206), Deflate() leaves INIT_STATE (:436-464)
		set_error("SetDictionary: not in the initial state"); // InvalidOperationException
		return B200Z_E_STATE;
	}
	if (len < 0 || (len > 0 && !dict)) {
		set_error("dictionary/count");
		return B200Z_E_ARG;
	}
	uint32_t a = 1;
	if (len > 0) {
		int rc = b200z_adler32(dict, len, &a);
		if (rc) return rc;
	}
	d->dict_adler = a;
	d->dict_set = true;
	if (len < kMinMatch) return B200Z_OK; // DeflaterEngine.cs:207-210: too short to matter, not even copied
	const int32_t keep = len > kMaxDist ? kMaxDist : len; // :212-216
	deflater_remember(d, dict + (len - keep), (size_t)keep, 2);
	return B200Z_OK;
}
int b200z_deflater_set_input(void *h, const uint8_t *buf, int32_t len) {
	DeflaterH *d = (DeflaterH *)h;
	if (d->finishing) {
		set_error("Finish() already called"); // Deflater.cs:335
		return B200Z_E_STATE;
	}
	if (len < 0 || (len > 0 && !buf)) {
		set_error("buffer/count");
		return B200Z_E_ARG;
	}
	if (len > 0) d->flushed_once = false;
	d->input.insert(d->input.end(), buf, buf + len);
	if (len > 0) {
		// engine.NeedsInput (DeflaterEngine.cs:187-190) is false until Deflate() has taken the bytes; a zero-length SetInput
		// leaves it true, so the stream classes do not even call Deflate() for it
		d->undrained = true;
		if (d->level <= 4) d->chunks.push_back(len); // the schedule matters for DeflateStored / DeflateFast only (trap T9)
	}
	return B200Z_OK;
}
int b200z_deflater_flush(void *h) {
	((DeflaterH *)h)->flushing = true;
	return B200Z_OK;
}
int b200z_deflater_finish(void *h) {
	DeflaterH *d = (DeflaterH *)h;
	d->flushing = d->finishing = true;
	return B200Z_OK;
}

int b200z_deflater_deflate(void *h, uint8_t *out, int32_t cap, int32_t *produced) {
	DeflaterH *d = (DeflaterH *)h;
	if (!produced || cap < 0 || (cap > 0 && !out)) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	*produced = 0;
	d->deflate_called = true;
	if (!d->flushing && !d->finishing) d->undrained = false; // BUSY_STATE: the engine takes everything SetInput handed over
	// make output due (Deflater.Deflate :427-522)
	if (!d->finished && d->pending_pos == d->pending.size() && (d->flushing || d->finishing)) {
		d->pending.clear();
		d->pending_pos = 0;
		if (!d->header_done && !d->raw) {
			uint8_t hd[2];
			zlib_header(d->level, hd, d->dict_set);
			d->pending.push_back(hd[0]);
			d->pending.push_back(hd[1]);
			if (d->dict_set) { // DICTID = Adler-32 of the whole dictionary, MSB first (:455-461)
				for (int sh = 24; sh >= 0; sh -= 8) d->pending.push_back((uint8_t)(d->dict_adler >> sh));
			}
		}
		d->header_done = true;
		if (d->finishing) {
			if (d->flushed_once) {
				// Flush() already emitted every block + the sync padding; Finish adds the final empty block:
				// levels 1-9 FlushBlock on an empty buffer = static header 011 + EOB 0000000 = value 3 in 10 bits
				// (DeflaterEngine.cs:750-768); level 0 an empty stored block 01 00 00 FF FF (:614-649)
				if (d->level == 0) {
					const uint8_t e5[5] = {1, 0, 0, 0xFF, 0xFF};
					d->pending.insert(d->pending.end(), e5, e5 + 5);
				} else {
					d->tail.put(d->pending, 3, 10);
					d->tail.align(d->pending);
				}
			} else {
				int rc = deflater_run_device(d, B200Z_END_FINISH);
				if (rc) return rc;
			}
			if (!d->raw) {
				d->pending.push_back((uint8_t)(d->adler >> 24));
				d->pending.push_back((uint8_t)(d->adler >> 16));
				d->pending.push_back((uint8_t)(d->adler >> 8));
				d->pending.push_back((uint8_t)d->adler);
			}
			d->finished = true;
		} else {
			// sync flush (level 0 skips the padding, :488)
			int rc = deflater_run_device(d, B200Z_END_FLUSH);
			if (rc) return rc;
			d->flushed_once = true;
			d->flushing = false;
		}
	}
	size_t avail = d->pending.size() - d->pending_pos;
	size_t take = avail < (size_t)cap ? avail : (size_t)cap;
	if (take) memcpy(out, d->pending.data() + d->pending_pos, take);
	d->pending_pos += take;
	d->total_out += (int64_t)take;
	*produced = (int32_t)take;
	return B200Z_OK;
}
int b200z_deflater_needs_input(void *h, int *flag) {
	// DeflaterEngine.NeedsInput (:187-190): false between SetInput and the Deflate() call that takes the bytes.  The
	// handle copies on SetInput, but it reports what the reference reports: DeflaterOutputStream.Write calls Deflate()
	// exactly while this is false (Streams/DeflaterOutputStream.cs:245-275), and whether that call happened before
	// Flush() / Finish() changes the bytes at levels 0-4 (b200z_history.undrained_last)
	*flag = ((DeflaterH *)h)->undrained ? 0 : 1;
	return B200Z_OK;
}
int b200z_deflater_is_finished(void *h, int *flag) {
	DeflaterH *d = (DeflaterH *)h;
	*flag = (d->finished && d->pending_pos == d->pending.size()) ? 1 : 0; // Deflater.cs:271-277
	return B200Z_OK;
}
int b200z_deflater_total_in(void *h, int64_t *v) {
	DeflaterH *d = (DeflaterH *)h;
	*v = d->total_in;
	return B200Z_OK;
}
int b200z_deflater_total_out(void *h, int64_t *v) {
	*v = ((DeflaterH *)h)->total_out;
	return B200Z_OK;
}
int b200z_deflater_adler(void *h, uint32_t *v) {
	DeflaterH *d = (DeflaterH *)h;
	*v = d->raw ? 0u : d->adler;
	return B200Z_OK;
}

// ---- Inflater handle ----------------------------------------------------------------------------------
// The reference's Inflater is a mode machine that stops at any bit and resumes (Inflater.cs:73-86, :429-552).  The kernel
// decodes whole plans, so the handle resumes at BLOCK granularity: k_inflate reports the last block header it reached
// (bit position, output position -- the "restart point"); when more input arrives decoding continues from that header
// with the last 32 KiB of output in front of it as the window image.  Bytes of a block that was only partly available
// are decoded again by the next run (they were already delivered; the new run's output replaces them byte for byte).
// Work is linear in the stream for any SetInput granularity as long as blocks are bounded (the reference's own Deflater
// cuts a block every 16384 symbols), and the handle holds one block of input / output plus the window, not the stream.
struct InflaterH {
	bool raw = false;
	std::vector<uint8_t> input; // compressed bytes from absolute offset in_base on (what lies in front of the restart point is dropped)
	int64_t in_base = 0, in_total = 0; // in_total = bytes handed over since Reset = in_base + input.size()
	bool new_input = false;
	bool finished = false;
	bool need_dict = false;
	bool header_done = false;
	uint32_t read_adler = 0;   // DICTID from the header (Inflater.readAdler)
	int64_t raw_off = 0;       // where the raw deflate data starts (behind the zlib header)
	int64_t rs_bit = 0;        // restart point: bit offset from raw_off of the block header decoding continues at
	int64_t rs_out = 0;        //                output position of that header
	std::vector<uint8_t> window; // last <= 32768 bytes of (dictionary ++ output[0, rs_out)): OutputWindow's contents there
	uint32_t run_adler = 1;    // Adler-32 of output[0, rs_out)
	std::vector<uint8_t> output; // decoded bytes from output position out_base on (delivered bytes in front of rs_out are dropped)
	int64_t out_base = 0;
	int64_t delivered = 0;
	int64_t consumed = 0; // bytes of input the decoder has used (header + raw + trailer), absolute
	uint32_t adler = 1;
	int error = 0;
	std::string error_msg;
	PinnedBuf hin, hout;
	DevBuf din, dout, dmeta;
	int64_t out_total() const { return out_base + (int64_t)output.size(); }
};

static const char *inflate_detail_msg(int detail) {
	switch (detail) {
	case 1: return "Unknown block type";
	case 2: return "broken uncompressed block";
	case 3: return "Illegal rep length code";
	case 4: return "Illegal rep dist code";
	case 5: return "Encountered invalid codelength 0";
	case 6: return "ValueOutOfRangeException: dynamic header code count";
	case 7: return "Cannot repeat previous code length when no other code length has been read";
	case 8: return "Cannot repeat code lengths past total number of data code lengths";
	case 9: return "Inflater dynamic header end-of-block code missing";
	case 10: return "Code lengths oversubscribed";
	case 11: return "Adler chksum doesn't match";
	case 12: return "GZIP crc sum mismatch";
	case 13: return "Number of bytes mismatch in footer";
	case 14: return "Error GZIP header, first magic byte doesn't match";
	case 15: return "Error GZIP header,  second magic byte doesn't match";
	case 16: return "Error GZIP header, data not in deflate format";
	case 17: return "Reserved flag bits in GZIP header != 0";
	case 18: return "Header CRC value mismatch";
	case 19: return "Header checksum illegal";
	case 20: return "Compression Method unknown";
	case 22: return "Needs a preset dictionary";
	default: return "corrupt deflate data";
	}
}

// decodes from the restart point to the end of what is available; fills output / consumed / finished and moves the
// restart point forward
static int inflater_run_device(InflaterH *d) {
	const int64_t slice_abs = d->raw_off + (d->rs_bit >> 3); // absolute offset of the first compressed byte of this run
	const int32_t sbit = (int32_t)(d->rs_bit & 7);
	const int64_t avail = d->in_total - slice_abs;
	if (avail < 0 || slice_abs < d->in_base) {
		set_error("inflater restart point outside the kept input");
		return B200Z_E_INTERNAL;
	}
	int64_t cap = avail * 8 + 65536;
	for (int attempt = 0; attempt < 8; attempt++) {
		b200z_plan *plan = nullptr;
		const int64_t D = (int64_t)d->window.size();
		int rc = b200z_inflate_plan_create_ex(1, &avail, &cap, B200Z_WRAP_RAW, D ? &D : nullptr, &plan);
		if (rc) return rc;
		if (sbit && (rc = b200z_inflate_plan_set_start_bits(plan, &sbit))) {
			b200z_plan_destroy(plan);
			return rc;
		}
		const uint8_t *inp = d->input.data() + (slice_abs - d->in_base);
		std::vector<uint8_t> slot;
		if (D) { // window image (dictionary, earlier output) directly in front of the compressed bytes
			slot.reserve((size_t)(D + avail));
			slot.insert(slot.end(), d->window.begin(), d->window.end());
			slot.insert(slot.end(), inp, inp + avail);
			inp = slot.data();
		}
		HostRunResult r;
		rc = run_plan_host(plan, &inp, d->hin, d->hout, d->din, d->dout, d->dmeta, r, false);
		int64_t rbit = 0, rout = 0;
		if (!rc) rc = b200z_plan_get_restart_points(plan, &rbit, &rout, nullptr);
		if (rc) {
			b200z_plan_destroy(plan);
			return rc;
		}
		const int st = r.status[0] & 0xFF, detail = (r.status[0] >> 8) & 0xFF;
		if (st == B200Z_E_NOMEM) {
			b200z_plan_destroy(plan);
			cap *= 8;
			continue;
		}
		// this run's output is output[rs_out, rs_out + out_len): it replaces what an earlier run decoded behind the
		// restart point (the same bytes, and at least as many)
		const uint8_t *o = d->hout.p + plan->out_off[0];
		d->output.resize((size_t)(d->rs_out - d->out_base));
		d->output.insert(d->output.end(), o, o + r.out_len[0]);
		b200z_plan_destroy(plan);
		if (st == B200Z_OK) {
			d->finished = true;
			d->consumed = slice_abs + r.in_used[0];
		} else if (st == B200Z_E_NEED_INPUT) {
			d->consumed = d->in_total;
		} else {
			d->error = st;
			d->error_msg = inflate_detail_msg(detail);
			return B200Z_OK;
		}
		if (rout < 0 || rout > r.out_len[0] || rbit < sbit || rbit > 8 * avail) {
			set_error("inflater restart point out of range");
			return B200Z_E_INTERNAL;
		}
		if (rout > 0 || rbit != sbit) {
			// move the restart point: checksum and window advance over output[rs_out, rs_out + rout)
			const uint8_t *seg = d->output.data() + (d->rs_out - d->out_base);
			if (!d->raw && rout > 0 && (rc = b200z_adler32(seg, rout, &d->run_adler))) return rc;
			if (rout >= 32768) d->window.assign(seg + rout - 32768, seg + rout);
			else {
				d->window.insert(d->window.end(), seg, seg + rout);
				if (d->window.size() > 32768) d->window.erase(d->window.begin(), d->window.end() - 32768);
			}
			d->rs_out += rout;
			d->rs_bit = (d->rs_bit & ~7ll) + rbit;
		}
		if (st == B200Z_E_NEED_INPUT) {
			// compressed bytes in front of the restart point are never looked at again
			const int64_t keep_from = d->raw_off + (d->rs_bit >> 3);
			if (keep_from > d->in_base) {
				d->input.erase(d->input.begin(), d->input.begin() + (keep_from - d->in_base));
				d->in_base = keep_from;
			}
		}
		return B200Z_OK;
	}
	set_error("output larger than any capacity tried");
	return B200Z_E_NOMEM;
}

// output in front of both the restart point and the delivery position is not needed any more
static void inflater_trim_output(InflaterH *d) {
	const int64_t keep_from = d->delivered < d->rs_out ? d->delivered : d->rs_out;
	if (keep_from - d->out_base >= 65536) {
		d->output.erase(d->output.begin(), d->output.begin() + (keep_from - d->out_base));
		d->out_base = keep_from;
	}
}

int b200z_inflater_create(int raw, void **h) {
	if (!h) {
		set_error("h");
		return B200Z_E_ARG;
	}
	InflaterH *d = new InflaterH();
	d->raw = raw != 0;
	*h = d;
	return B200Z_OK;
}
int b200z_inflater_destroy(void *h) {
	delete (InflaterH *)h;
	return B200Z_OK;
}
int b200z_inflater_reset(void *h) {
	InflaterH *d = (InflaterH *)h;
	d->input.clear();
	d->output.clear();
	d->window.clear();
	d->new_input = d->finished = d->need_dict = d->header_done = false;
	d->in_base = d->in_total = 0;
	d->raw_off = 0;
	d->rs_bit = d->rs_out = 0;
	d->run_adler = 1;
	d->out_base = 0;
	d->delivered = 0;
	d->consumed = 0;
	d->adler = 1;
	d->error = 0;
	d->read_adler = 0;
	return B200Z_OK;
}
int b200z_inflater_set_dictionary(void *h, const uint8_t *dict, int32_t len) {
	InflaterH *d = (InflaterH *)h;
	// Inflater.SetDictionary :589-620
	if (len < 0 || (len > 0 && !dict)) {
		set_error("buffer/count");
		return B200Z_E_ARG;
	}
	if (!d->need_dict) {
		set_error("Dictionary is not needed"); // InvalidOperationException
		return B200Z_E_STATE;
	}
	uint32_t a = 1;
	if (len > 0) {
		int rc = b200z_adler32(dict, len, &a);
		if (rc) return rc;
	}
	if (a != d->read_adler) {
		set_error("Wrong adler checksum"); // SharpZipBaseException
		return B200Z_E_DATA;
	}
	const int32_t keep = len > 32768 ? 32768 : len; // OutputWindow.CopyDict :160-166
	d->window.assign(dict + (len - keep), dict + len);
	d->need_dict = false;
	d->adler = 1;
	d->new_input = true; // what was handed over behind the header can be decoded now
	return B200Z_OK;
}
int b200z_inflater_set_input(void *h, const uint8_t *buf, int32_t len) {
	InflaterH *d = (InflaterH *)h;
	if (len < 0 || (len > 0 && !buf)) {
		set_error("buffer/count");
		return B200Z_E_ARG;
	}
	if (d->in_total > d->consumed && !d->finished) {
		set_error("Old input was not completely processed"); // StreamManipulator.cs:263
		return B200Z_E_STATE;
	}
	if (d->finished) {
		// after the end of the stream the reference keeps unread bytes as RemainingInput; replace them
		d->input.resize((size_t)(d->consumed - d->in_base));
		d->in_total = d->consumed;
	}
	d->input.insert(d->input.end(), buf, buf + len);
	d->in_total += len;
	d->new_input = true;
	return B200Z_OK;
}

int b200z_inflater_inflate(void *h, uint8_t *out, int32_t cap, int32_t *produced) {
	InflaterH *d = (InflaterH *)h;
	if (!produced || cap < 0 || (cap > 0 && !out)) {
		set_error("bad arguments");
		return B200Z_E_ARG;
	}
	*produced = 0;
	if (d->error) {
		set_error("%s", d->error_msg.c_str());
		return d->error;
	}
	if (!d->finished && d->new_input) {
		d->new_input = false;
		if (!d->raw && !d->header_done) {
			// (nothing is dropped from `input` before the header is done: in_base == 0 here)
			if (d->input.size() < 2) {
				d->consumed = (int64_t)d->input.size();
				return B200Z_OK; // DecodeHeader needs 16 bits (Inflater.cs:209-215)
			}
			const int header = (d->input[0] << 8) | d->input[1];
			if (header % 31 != 0) {
				d->error = B200Z_E_DATA;
				d->error_msg = "Header checksum illegal";
			} else if ((header & 0x0f00) != (8 << 8)) {
				d->error = B200Z_E_DATA;
				d->error_msg = "Compression Method unknown";
			}
			if (d->error) {
				set_error("%s", d->error_msg.c_str());
				return d->error;
			}
			if (header & 0x0020) {
				// PRESET_DICT: DecodeDict reads the 4-byte DICTID MSB first (:185-203); Inflate() then returns 0 until
				// SetDictionary supplies a dictionary with that Adler-32
				if (d->input.size() < 6) {
					d->consumed = (int64_t)d->input.size();
					return B200Z_OK;
				}
				d->read_adler = ((uint32_t)d->input[2] << 24) | ((uint32_t)d->input[3] << 16) | ((uint32_t)d->input[4] << 8) | d->input[5];
				d->need_dict = true;
				d->adler = d->read_adler; // Inflater.Adler reports readAdler while the dictionary is awaited (:823-836)
				d->header_done = true;
				d->raw_off = 6;
				d->consumed = 6;
				d->new_input = true; // the bytes behind the DICTID wait for SetDictionary
				return B200Z_OK;
			}
			d->header_done = true;
			d->raw_off = 2;
		}
		if (d->need_dict) {
			d->new_input = true;
			return B200Z_OK;
		}
		int rc = inflater_run_device(d);
		if (rc) return rc;
		if (d->finished && !d->raw) {
			// Adler32 trailer, read MSB first (DecodeChksum :397-418); until it is complete the reference withholds
			// nothing that was already decoded but is not "finished"
			if (d->in_total - d->consumed < 4) {
				d->finished = false;
				d->consumed = d->in_total;
				// keep output; the last block is decoded again when the trailer arrives
			} else {
				const uint8_t *t = d->input.data() + (d->consumed - d->in_base);
				const uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
				uint32_t got = d->run_adler; // output[0, rs_out) is in it already
				const int64_t rest = d->out_total() - d->rs_out;
				int rc2 = rest > 0 ? b200z_adler32(d->output.data() + (d->rs_out - d->out_base), rest, &got) : B200Z_OK;
				if (rc2) return rc2;
				d->adler = got;
				if (got != want) {
					d->error = B200Z_E_DATA;
					d->error_msg = "Adler chksum doesn't match";
				}
				d->consumed += 4;
			}
		}
	}
	size_t avail = (size_t)(d->out_total() - d->delivered);
	size_t take = avail < (size_t)cap ? avail : (size_t)cap;
	if (take) memcpy(out, d->output.data() + (d->delivered - d->out_base), take);
	d->delivered += (int64_t)take;
	*produced = (int32_t)take;
	inflater_trim_output(d);
	if (take == 0 && d->error) {
		set_error("%s", d->error_msg.c_str());
		return d->error;
	}
	return B200Z_OK;
}
int b200z_inflater_needs_input(void *h, int *flag) {
	InflaterH *d = (InflaterH *)h;
	*flag = (d->in_total <= d->consumed) ? 1 : 0; // StreamManipulator.IsNeedingInput
	return B200Z_OK;
}
int b200z_inflater_needs_dictionary(void *h, int *flag) {
	*flag = ((InflaterH *)h)->need_dict ? 1 : 0; // mode == DECODE_DICT && neededBits == 0 (Inflater.cs:792-798)
	return B200Z_OK;
}
int b200z_inflater_is_finished(void *h, int *flag) {
	InflaterH *d = (InflaterH *)h;
	// Inflater.cs:806-812; a trailer that did not verify is an exception from Inflate() in the reference (DecodeChksum
	// :397-418), so the stream is not "finished" as long as that error has not been delivered
	*flag = (d->finished && !d->error && d->delivered == d->out_total()) ? 1 : 0;
	return B200Z_OK;
}
int b200z_inflater_remaining_input(void *h, int32_t *v) {
	InflaterH *d = (InflaterH *)h;
	*v = (int32_t)(d->in_total - d->consumed);
	return B200Z_OK;
}
int b200z_inflater_total_in(void *h, int64_t *v) {
	*v = ((InflaterH *)h)->consumed; // totalIn - RemainingInput (Inflater.cs:862-868)
	return B200Z_OK;
}
int b200z_inflater_total_out(void *h, int64_t *v) {
	*v = ((InflaterH *)h)->delivered;
	return B200Z_OK;
}
int b200z_inflater_adler(void *h, uint32_t *v) {
	InflaterH *d = (InflaterH *)h;
	*v = d->raw ? 0u : d->adler;
	return B200Z_OK;
}

} // extern "C"
