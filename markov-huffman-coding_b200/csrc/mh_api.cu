// mh_api.cu — the C ABI of libmh_gpu.so (include/mh_gpu.h): handles, workspaces, the device entry points and the
// host-buffer session that stands in for i_coding_provider::compress / decompress (reference src/coding.cpp:61-160)
// and for the table construction in main (src/main.cpp:164-183).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "mh_host.hpp"
#include "mh_internal.hpp"

static std::atomic<uint64_t> g_table_serial{0};

struct mh_table {
	mh::CodingTable impl;
	const uint64_t serial = ++g_table_serial;   // never reused: identifies the table the device images were flattened from
};

namespace mh {

std::atomic<uint64_t> g_kernel_launches{0};
static thread_local std::string t_last_error;

void set_last_error(const char* msg) { t_last_error = msg ? msg : ""; }
int table_order(const mh_table* t) { return t->impl.order; }
uint64_t table_serial(const mh_table* t) { return t->serial; }

int cuda_fail(cudaError_t e, const char* what) {
	t_last_error = std::string(what) + ": " + cudaGetErrorString(e);
	cudaGetLastError();   // clear the sticky-free error state
	return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MH_ERR_NO_DEVICE : MH_ERR_CUDA;
}

namespace {
struct ProfRecord { const char* name; cudaEvent_t a, b; };
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;   // guards g_prof: launches may come from one host thread per GPU
std::vector<ProfRecord> g_prof;
}  // namespace

ProfScope::ProfScope(const char* name, cudaStream_t s) : slot(-1), st(s) {
	if(!g_prof_on.load(std::memory_order_relaxed)) return;
	ProfRecord r{name, nullptr, nullptr};
	if(cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
	cudaEventRecord(r.a, st);
	std::lock_guard<std::mutex> lock(g_prof_mu);
	g_prof.push_back(r);
	slot = int(g_prof.size()) - 1;
	end = r.b;
}
ProfScope::~ProfScope() {
	if(slot >= 0) cudaEventRecord(end, st);
}

namespace {
constexpr int kMaxDevices = 64;
int current_device_slot() {
	int dev = 0;
	if(cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
	return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
std::atomic<int> g_sm_count[kMaxDevices];
std::atomic<int> g_smem_optin[kMaxDevices];
}  // namespace

int sm_count() {
	const int dev = current_device_slot();
	int v = g_sm_count[dev].load(std::memory_order_relaxed);
	if(!v) {
		if(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 148; }
		g_sm_count[dev].store(v, std::memory_order_relaxed);
	}
	return v;
}

int max_smem_optin() {
	const int dev = current_device_slot();
	int v = g_smem_optin[dev].load(std::memory_order_relaxed);
	if(!v) {
		if(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 48 * 1024; }
		g_smem_optin[dev].store(v, std::memory_order_relaxed);
	}
	return v;
}

bool first_use_on_device(std::atomic<uint64_t>& done_mask) {
	const uint64_t bit = 1ull << current_device_slot();
	return (done_mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

// ---- tunables -------------------------------------------------------------------------------------------
namespace {
struct TunableDef { const char* name; const char* env; };
const TunableDef kTunables[kTunCount] = {
    {"enc_fmt", "MH_ENC_FMT"},
    {"dec_sub_bits_markov", "MH_DEC_SUB_BITS_MARKOV"},
    {"dec_sub_bits_huffman", "MH_DEC_SUB_BITS_HUFFMAN"},
    {"dec_pair", "MH_DEC_PAIR"},
    {"dec_write_threads", "MH_DEC_WRITE_THREADS"},
    {"pipe_min_bytes", "MH_PIPE_MIN_BYTES"},
    {"pipe_chunk_bytes", "MH_PIPE_CHUNK_BYTES"},
    {"enc_pipe_chunk_bytes", "MH_ENC_PIPE_CHUNK_BYTES"},
    {"enc_tma", "MH_ENC_TMA"},
    {"dec_cp_geo", "MH_DEC_CP_GEO"},
    {"enc_warp", "MH_ENC_WARP"},
    {"enc_spt", "MH_ENC_SPT"},
};
std::atomic<long long> g_tunables[kTunCount];
struct TunableInit {
	TunableInit() {
		for(int i = 0; i < kTunCount; ++i) {
			const char* e = getenv(kTunables[i].env);   // the only getenv calls of the library: once, at load time
			g_tunables[i].store(e && *e ? strtoll(e, nullptr, 10) : -1, std::memory_order_relaxed);
		}
	}
} g_tunable_init;
}  // namespace

long long tunable(Tunable t) { return g_tunables[t].load(std::memory_order_relaxed); }

}  // namespace mh

using namespace mh;

extern "C" {

const char* mh_status_string(int status) {
	switch(status) {
		case MH_OK: return "ok";
		case MH_ERR_INVALID_ARG: return "invalid argument";
		case MH_ERR_CUDA: return "CUDA error";
		case MH_ERR_NO_DEVICE: return "no CUDA device";
		case MH_ERR_CAPACITY: return "output buffer too small";
		case MH_ERR_BAD_TABLE: return "encoding table is truncated or malformed";
		case MH_ERR_CODE_TOO_LONG: return "a codeword exceeds 56 bits";
		case MH_ERR_BAD_HEADER: return "Input appears corrupt";
		case MH_ERR_TYPE_MISMATCH: return "File encoding method does not match provided encoding table";
		case MH_ERR_CORRUPT_STREAM: return "compressed stream is corrupt";
		case MH_ERR_COUNT_WRAPPED: return "a symbol count is a non-zero multiple of 2^32";
		case MH_ERR_NOT_CONVERGED: return "decoder seams did not converge";
		case MH_ERR_WORKSPACE: return "workspace too small";
	}
	return "unknown status";
}

const char* mh_last_error(void) { return t_last_error.c_str(); }

int mh_device_memory(int device, uint64_t* free_bytes, uint64_t* total_bytes) {
	if(!free_bytes || !total_bytes) return MH_ERR_INVALID_ARG;
	int ndev = 0;
	if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); t_last_error = "no usable CUDA device"; return MH_ERR_NO_DEVICE; }
	if(device < 0 || device >= ndev) return MH_ERR_INVALID_ARG;
	MH_CUDA(cudaSetDevice(device));
	size_t f = 0, t = 0;
	MH_CUDA(cudaMemGetInfo(&f, &t));
	*free_bytes = f;
	*total_bytes = t;
	return MH_OK;
}

int mh_device_count(void) {
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int mh_version(void) { return 100; }

uint64_t mh_kernel_launches(void) { return g_kernel_launches.load(); }

int mh_profile_enable(int on) {
	std::lock_guard<std::mutex> lock(g_prof_mu);
	for(auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
	g_prof.clear();
	g_prof_on = on != 0;
	return MH_OK;
}

int mh_profile_report(char* out, size_t cap, size_t* n_out) {
	if(!n_out) return MH_ERR_INVALID_ARG;
	struct Agg { const char* name; uint64_t launches; double ms; };
	std::vector<Agg> agg;
	std::vector<ProfRecord> recs;
	{
		std::lock_guard<std::mutex> lock(g_prof_mu);
		recs.swap(g_prof);
	}
	for(auto& r : recs) {
		float ms = 0.f;
		const bool ok = cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess;
		cudaEventDestroy(r.a);
		cudaEventDestroy(r.b);
		if(!ok) { cudaGetLastError(); continue; }
		size_t k = 0;
		while(k < agg.size() && strcmp(agg[k].name, r.name) != 0) ++k;
		if(k == agg.size()) agg.push_back({r.name, 0, 0.0});
		agg[k].launches += 1;
		agg[k].ms += ms;
	}
	std::string js = "{";
	for(size_t k = 0; k < agg.size(); ++k) {
		char buf[256];
		snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f}", k ? ", " : "", agg[k].name,
		         (unsigned long long) agg[k].launches, agg[k].ms);
		js += buf;
	}
	js += "}";
	*n_out = js.size() + 1;
	if(js.size() + 1 > cap) return MH_ERR_CAPACITY;
	memcpy(out, js.c_str(), js.size() + 1);
	return MH_OK;
}

int mh_tunable_set(const char* name, long long value) {
	if(!name) return MH_ERR_INVALID_ARG;
	for(int i = 0; i < kTunCount; ++i)
		if(strcmp(name, kTunables[i].name) == 0) { g_tunables[i].store(value, std::memory_order_relaxed); return MH_OK; }
	return MH_ERR_INVALID_ARG;
}

int mh_tunable_get(const char* name, long long* value) {
	if(!name || !value) return MH_ERR_INVALID_ARG;
	for(int i = 0; i < kTunCount; ++i)
		if(strcmp(name, kTunables[i].name) == 0) { *value = g_tunables[i].load(std::memory_order_relaxed); return MH_OK; }
	return MH_ERR_INVALID_ARG;
}

// ---------------------------------------------------------------------------------------------------------
// host tables
// ---------------------------------------------------------------------------------------------------------
int mh_table_from_counts(const uint64_t* counts, int order, mh_table** out) {
	if(!counts || !out) return MH_ERR_INVALID_ARG;
	mh_table* t = new(std::nothrow) mh_table;
	if(!t) return MH_ERR_INVALID_ARG;
	int rc = CodingTable::from_counts(counts, order, t->impl);
	if(rc != MH_OK) { delete t; return rc; }
	*out = t;
	return MH_OK;
}

int mh_table_from_bytes(const uint8_t* bytes, size_t n, mh_table** out) {
	if(!out) return MH_ERR_INVALID_ARG;
	mh_table* t = new(std::nothrow) mh_table;
	if(!t) return MH_ERR_INVALID_ARG;
	int rc = CodingTable::from_bytes(bytes, n, t->impl);
	if(rc != MH_OK) { delete t; return rc; }
	*out = t;
	return MH_OK;
}

int mh_table_serialize(const mh_table* t, uint8_t* out, size_t cap, size_t* n_out) {
	if(!t || !n_out) return MH_ERR_INVALID_ARG;
	std::vector<uint8_t> v = t->impl.serialize();
	*n_out = v.size();
	if(v.size() > cap) return MH_ERR_CAPACITY;
	if(!v.empty()) memcpy(out, v.data(), v.size());
	return MH_OK;
}

int mh_table_order(const mh_table* t) { return t ? t->impl.order : MH_ERR_INVALID_ARG; }

int mh_table_context_empty(const mh_table* t, int prev) {
	if(!t || prev < 0 || prev > 255) return MH_ERR_INVALID_ARG;
	return t->impl.tree_for(prev).empty() ? 1 : 0;
}

int mh_table_code(const mh_table* t, int prev, int c, uint8_t bits[32], int* len) {
	if(!t || !len || prev < 0 || prev > 255 || c < 0 || c > 255) return MH_ERR_INVALID_ARG;
	const Codeword& cw = t->impl.tree_for(prev).code(c);
	*len = cw.length;
	if(bits) memcpy(bits, cw.bytes.data(), 32);
	return MH_OK;
}

int mh_table_max_code_bits(const mh_table* t) { return t ? t->impl.max_code_bits() : MH_ERR_INVALID_ARG; }

int mh_table_code_lengths(const mh_table* t, uint8_t* lens, size_t cap) {
	if(!t || !lens) return MH_ERR_INVALID_ARG;
	const size_t ntab = t->impl.trees.size();
	if(cap < ntab * 256) return MH_ERR_CAPACITY;
	memset(lens, 0, ntab * 256);
	for(size_t k = 0; k < ntab; ++k) {
		if(t->impl.trees[k].empty()) continue;
		for(int c = 0; c < 256; ++c) lens[k * 256 + c] = uint8_t(t->impl.trees[k].code(c).length);
	}
	return MH_OK;
}

int mh_table_lookup(const mh_table* t, int prev, int window, int* kind, int* value, int* depth) {
	if(!t || prev < 0 || prev > 255 || window < 0 || window > 255 || !kind || !value || !depth) return MH_ERR_INVALID_ARG;
	const CodeTree& tr = t->impl.tree_for(prev);
	const int n = tr.lut(window);
	if(n == kNoChild) { *kind = 0; *value = 0; *depth = 0; return MH_OK; }
	*kind = tr.nodes[n].internal ? 2 : 1;
	*value = tr.nodes[n].symbol;
	*depth = tr.nodes[n].depth;
	return MH_OK;
}

int mh_table_pair_lut(const mh_table* t, uint32_t* table, uint8_t* maps, uint32_t* rows, uint32_t* ctx_rows) {
	if(!t || !table || !maps || !rows || !ctx_rows) return MH_ERR_INVALID_ARG;
	*ctx_rows = 0;
	*rows = t->impl.flatten_pairlut(table, maps, kPairMaxRows, ctx_rows);
	return MH_OK;
}

int mh_table_debug_dump(const mh_table* t, char* out, size_t cap, size_t* n_out) {
	if(!t || !n_out) return MH_ERR_INVALID_ARG;
	std::string s = t->impl.debug_dump();
	*n_out = s.size();
	if(s.size() > cap) return MH_ERR_CAPACITY;
	if(!s.empty()) memcpy(out, s.data(), s.size());
	return MH_OK;
}

void mh_table_destroy(mh_table* t) { delete t; }

// ---------------------------------------------------------------------------------------------------------
// device tables and scratch
// ---------------------------------------------------------------------------------------------------------
// Uploads are stream-ordered and never block the host: the flat image is built in a pinned buffer owned by the
// handle; an event guards that buffer against being rewritten while a previous copy is still reading it.
static int upload_codebook(const mh_table* t, mh_codebook* cb, cudaStream_t st) {
	if(!cb->d_enc) MH_CUDA(cudaMalloc(&cb->d_enc, 65536 * sizeof(uint64_t)));
	if(!cb->h_stage) MH_CUDA(cudaMallocHost(&cb->h_stage, 65536 * sizeof(uint64_t)));
	if(!cb->uploaded) MH_CUDA(cudaEventCreateWithFlags(&cb->uploaded, cudaEventDisableTiming));
	else MH_CUDA(cudaEventSynchronize(cb->uploaded));
	int rc = t->impl.flatten_codebook(cb->h_stage);
	if(rc != MH_OK) return rc;
	MH_CUDA(cudaMemcpyAsync(cb->d_enc, cb->h_stage, t->impl.trees.size() * 256 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
	cb->order = t->impl.order;
	cb->max_bits = t->impl.max_code_bits();
	cb->device_built = false;
	// preferred encoder table: rows for the live contexts only (text: a few dozen), next row carried by the entry
	cb->ctx_rows = 0;
	if(cb->max_bits <= kEncCtxMaxBits) {
		if(!cb->d_ctx) MH_CUDA(cudaMalloc(&cb->d_ctx, size_t(kEncCtxMaxRows) * 256 * sizeof(uint32_t)));
		if(!cb->h_ctx) MH_CUDA(cudaMallocHost(&cb->h_ctx, size_t(kEncCtxMaxRows) * 256 * sizeof(uint32_t)));
		cb->ctx_rows = t->impl.flatten_ctx(cb->h_ctx, kEncCtxMaxRows);
		if(cb->ctx_rows) MH_CUDA(cudaMemcpyAsync(cb->d_ctx, cb->h_ctx, size_t(cb->ctx_rows) * 256 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
	}
	// the box table is only needed when the context rows are not available (or a test forces another format)
	cb->has_box = cb->max_bits <= kEncBoxMaxBits && (cb->ctx_rows == 0 || tunable(kTunEncFmt) >= 0);
	if(cb->has_box) {
		if(!cb->d_box) MH_CUDA(cudaMalloc(&cb->d_box, 257 * 257 * sizeof(uint32_t)));
		if(!cb->h_box) MH_CUDA(cudaMallocHost(&cb->h_box, 257 * 257 * sizeof(uint32_t)));
		t->impl.live_range(cb->box_lo, cb->box_r);
		t->impl.flatten_box(cb->box_lo, cb->box_r, cb->h_box);
		const size_t entries = cb->order ? size_t(cb->box_r + 1) * (cb->box_r + 1) : 256;
		MH_CUDA(cudaMemcpyAsync(cb->d_box, cb->h_box, entries * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
	}
	MH_CUDA(cudaEventRecord(cb->uploaded, st));
	return MH_OK;
}

static int upload_dectable(const mh_table* t, mh_dectable* dt, cudaStream_t st) {
	if(!dt->d_lut) MH_CUDA(cudaMalloc(&dt->d_lut, 65536 * sizeof(uint16_t)));
	// one allocation: walk table [ntab * 512] u32, then the second-level rows [kExtRows * 256] u16
	const size_t ext_bytes = size_t(kExtRows) * 256 * sizeof(uint16_t);
	if(!dt->d_walk) MH_CUDA(cudaMalloc(&dt->d_walk, 256 * 512 * sizeof(uint32_t) + ext_bytes));
	if(!dt->h_lut) MH_CUDA(cudaMallocHost(&dt->h_lut, 65536 * sizeof(uint16_t)));
	if(!dt->h_walk) MH_CUDA(cudaMallocHost(&dt->h_walk, 256 * 512 * sizeof(uint32_t) + ext_bytes));
	if(!dt->uploaded) MH_CUDA(cudaEventCreateWithFlags(&dt->uploaded, cudaEventDisableTiming));
	else MH_CUDA(cudaEventSynchronize(dt->uploaded));
	const size_t ntab = t->impl.trees.size();
	const uint32_t ext_rows = t->impl.flatten_dectable(dt->h_lut, dt->h_walk, reinterpret_cast<uint16_t*>(dt->h_walk + ntab * 512));
	MH_CUDA(cudaMemcpyAsync(dt->d_lut, dt->h_lut, ntab * 256 * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
	MH_CUDA(cudaMemcpyAsync(dt->d_walk, dt->h_walk, ntab * 512 * sizeof(uint32_t) + size_t(ext_rows) * 256 * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
	// two-symbol table over the live contexts (text: a few dozen rows), preferred by the decoder when it exists
	if(!dt->d_pair) MH_CUDA(cudaMalloc(&dt->d_pair, kDecPairBytes));
	if(!dt->h_pair) MH_CUDA(cudaMallocHost(&dt->h_pair, kDecPairBytes));
	{
		static thread_local uint8_t maps[256 + kPairMaxRows * 257];
		dt->pair_rows = t->impl.flatten_pairlut(dt->h_pair, maps, kPairMaxRows, &dt->pair_ctx_rows);
		if(dt->pair_rows) {
			const size_t map_bytes = 256 + kPairMaxRows + size_t(dt->pair_ctx_rows) * 256;
			memcpy(dt->h_pair + size_t(dt->pair_rows) * 256, maps, map_bytes);
			MH_CUDA(cudaMemcpyAsync(dt->d_pair, dt->h_pair, size_t(dt->pair_rows) * 1024 + map_bytes, cudaMemcpyHostToDevice, st));
		}
	}
	MH_CUDA(cudaEventRecord(dt->uploaded, st));
	dt->order = t->impl.order;
	dt->max_bits = t->impl.max_code_bits();
	return MH_OK;
}

static void release_book(mh_codebook* cb) {
	if(cb->uploaded) { cudaEventSynchronize(cb->uploaded); cudaEventDestroy(cb->uploaded); }
	if(cb->d_enc) cudaFree(cb->d_enc);
	if(cb->h_stage) cudaFreeHost(cb->h_stage);
	if(cb->d_box) cudaFree(cb->d_box);
	if(cb->h_box) cudaFreeHost(cb->h_box);
	if(cb->d_ctx) cudaFree(cb->d_ctx);
	if(cb->h_ctx) cudaFreeHost(cb->h_ctx);
	if(cb->d_meta) cudaFree(cb->d_meta);
	*cb = mh_codebook();
}

static void release_dec(mh_dectable* dt) {
	if(dt->uploaded) { cudaEventSynchronize(dt->uploaded); cudaEventDestroy(dt->uploaded); }
	if(dt->d_lut) cudaFree(dt->d_lut);
	if(dt->d_walk) cudaFree(dt->d_walk);
	if(dt->h_lut) cudaFreeHost(dt->h_lut);
	if(dt->h_walk) cudaFreeHost(dt->h_walk);
	if(dt->d_pair) cudaFree(dt->d_pair);
	if(dt->h_pair) cudaFreeHost(dt->h_pair);
	*dt = mh_dectable();
}

}  // extern "C"
namespace mh {
int upload_codebook_for(const mh_table* t, mh_codebook* cb, cudaStream_t st) { return upload_codebook(t, cb, st); }
int upload_dectable_for(const mh_table* t, mh_dectable* dt, cudaStream_t st) { return upload_dectable(t, dt, st); }
void release_codebook(mh_codebook* cb) { release_book(cb); }
void release_dectable(mh_dectable* dt) { release_dec(dt); }
}  // namespace mh
extern "C" {

int mh_codebook_create(const mh_table* t, mh_codebook** out) {
	if(!t || !out) return MH_ERR_INVALID_ARG;
	mh_codebook* cb = new(std::nothrow) mh_codebook;
	if(!cb) return MH_ERR_INVALID_ARG;
	int rc = upload_codebook(t, cb, nullptr);
	if(rc != MH_OK) { mh_codebook_destroy(cb); return rc; }
	*out = cb;
	return MH_OK;
}

int mh_codebook_update(mh_codebook* cb, const mh_table* t, mh_stream_t stream) {
	if(!cb || !t) return MH_ERR_INVALID_ARG;
	return upload_codebook(t, cb, static_cast<cudaStream_t>(stream));
}

void mh_codebook_destroy(mh_codebook* cb) {
	if(!cb) return;
	release_book(cb);
	delete cb;
}

int mh_codebook_create_empty(mh_codebook** out) {
	if(!out) return MH_ERR_INVALID_ARG;
	mh_codebook* cb = new(std::nothrow) mh_codebook;
	if(!cb) return MH_ERR_INVALID_ARG;
	*out = cb;
	return MH_OK;
}

int mh_codebook_build_device(mh_codebook* cb, const uint64_t* d_counts, int order, mh_stream_t stream) {
	return launch_build_codebook(reinterpret_cast<const unsigned long long*>(d_counts), order, cb, static_cast<cudaStream_t>(stream));
}

int mh_codebook_download(mh_codebook* cb, uint64_t* enc, uint32_t* ctx, uint32_t* meta) {
	if(!cb || !cb->d_enc) return MH_ERR_INVALID_ARG;
	MH_CUDA(cudaDeviceSynchronize());
	const size_t ntab = cb->order ? 256 : 1;
	if(enc) MH_CUDA(cudaMemcpy(enc, cb->d_enc, ntab * 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
	uint32_t m[8] = {cb->ctx_rows, 0, uint32_t(cb->max_bits), 0, 0, 0, 0, 0};
	if(cb->device_built) MH_CUDA(cudaMemcpy(m, cb->d_meta, sizeof m, cudaMemcpyDeviceToHost));
	if(meta) memcpy(meta, m, sizeof m);
	if(ctx && cb->d_ctx && m[0] && m[0] <= uint32_t(kEncCtxMaxRows)) MH_CUDA(cudaMemcpy(ctx, cb->d_ctx, size_t(m[0]) * 256 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return MH_OK;
}

int mh_dectable_create(const mh_table* t, mh_dectable** out) {
	if(!t || !out) return MH_ERR_INVALID_ARG;
	mh_dectable* dt = new(std::nothrow) mh_dectable;
	if(!dt) return MH_ERR_INVALID_ARG;
	int rc = upload_dectable(t, dt, nullptr);
	if(rc != MH_OK) { mh_dectable_destroy(dt); return rc; }
	*out = dt;
	return MH_OK;
}

int mh_dectable_update(mh_dectable* dt, const mh_table* t, mh_stream_t stream) {
	if(!dt || !t) return MH_ERR_INVALID_ARG;
	return upload_dectable(t, dt, static_cast<cudaStream_t>(stream));
}

void mh_dectable_destroy(mh_dectable* dt) {
	if(!dt) return;
	release_dec(dt);
	delete dt;
}

int mh_workspace_create(uint64_t max_input_bytes, uint64_t max_payload_bytes, mh_workspace** out) {
	if(!out) return MH_ERR_INVALID_ARG;
	mh_workspace* ws = new(std::nothrow) mh_workspace;
	if(!ws) return MH_ERR_INVALID_ARG;
	auto fail = [&](int rc) { mh_workspace_destroy(ws); return rc; };
#define WS_CUDA(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) return fail(cuda_fail(e_, #call)); } while(0)
	WS_CUDA(cudaMalloc(&ws->counters, 16 * sizeof(uint32_t)));
	WS_CUDA(cudaMalloc(&ws->hist_params, 8 * sizeof(uint32_t)));
	WS_CUDA(cudaMemset(ws->counters, 0, 16 * sizeof(uint32_t)));
	WS_CUDA(cudaMemset(ws->hist_params, 0, 8 * sizeof(uint32_t)));
	ws->enc_tiles_cap = encode_tiles_for(max_input_bytes) + 1;
	WS_CUDA(cudaMalloc(&ws->enc_desc, ws->enc_tiles_cap * (2 * sizeof(uint64_t) + sizeof(uint32_t)) + 4096));   // 20 bytes per tile cover either kernel's descriptors (K2w: 16 per tile + 136 per 64 tiles)
	ws->dec_subs_cap = decode_max_subs(max_payload_bytes);
	ws->dec_chunks_cap = ws->dec_subs_cap / (kDecThreads - kDecWarmSubs) + 2;
	if(ws->dec_chunks_cap < 1024) ws->dec_chunks_cap = 1024;   // short streams are cut into one chunk per SM
	WS_CUDA(cudaMalloc(&ws->dec_state, ws->dec_subs_cap * sizeof(uint32_t)));
	WS_CUDA(cudaMalloc(&ws->dec_count, ws->dec_subs_cap * sizeof(uint32_t)));
	WS_CUDA(cudaMalloc(&ws->dec_prefix, ws->dec_subs_cap * sizeof(uint32_t)));
	WS_CUDA(cudaMalloc(&ws->dec_seam, ws->dec_chunks_cap * sizeof(uint32_t)));
	WS_CUDA(cudaMalloc(&ws->dec_chunk_total, ws->dec_chunks_cap * sizeof(uint64_t)));
	WS_CUDA(cudaMalloc(&ws->dec_chunk_base, (ws->dec_chunks_cap + 1) * sizeof(uint64_t)));
	WS_CUDA(cudaMalloc(&ws->dec_flags, 8 * sizeof(uint32_t)));
#undef WS_CUDA
	*out = ws;
	return MH_OK;
}

void mh_workspace_destroy(mh_workspace* ws) {
	if(!ws) return;
	void* ptrs[] = {ws->enc_desc, ws->counters, ws->hist_params, ws->dec_state,
	                ws->dec_count, ws->dec_prefix, ws->dec_seam, ws->dec_chunk_total, ws->dec_chunk_base, ws->dec_flags};
	for(void* p : ptrs)
		if(p) cudaFree(p);
	delete ws;
}

// ---------------------------------------------------------------------------------------------------------
// the three device entry points
// ---------------------------------------------------------------------------------------------------------
int mh_gpu_histogram(const uint8_t* d_in, uint64_t n, uint8_t prev0, int order, uint64_t* d_counts, mh_workspace* ws,
                     mh_stream_t stream) {
	return launch_histogram(d_in, n, prev0, order, reinterpret_cast<unsigned long long*>(d_counts), ws, static_cast<cudaStream_t>(stream));
}

int mh_gpu_encode(const uint8_t* d_in, uint64_t n, uint8_t prev0, const mh_codebook* cb, uint64_t bit_base, uint8_t* d_out,
                  uint64_t out_capacity, uint64_t* d_result, mh_workspace* ws, mh_stream_t stream) {
	return launch_encode(d_in, n, prev0, cb, bit_base, d_out, out_capacity, reinterpret_cast<unsigned long long*>(d_result), ws,
	                     static_cast<cudaStream_t>(stream));
}

int mh_gpu_decode(const uint8_t* d_bits, uint64_t bit_base, uint64_t n_bits, uint8_t prev0, const mh_dectable* dt, uint8_t* d_out,
                  uint64_t out_capacity, uint64_t* d_result, mh_workspace* ws, mh_stream_t stream) {
	return launch_decode(d_bits, bit_base, n_bits, prev0, dt, d_out, out_capacity, reinterpret_cast<unsigned long long*>(d_result), ws,
	                     static_cast<cudaStream_t>(stream), 2);
}

int mh_gpu_decode_shard(const uint8_t* d_bits, uint32_t start_bit, uint64_t n_bits, uint64_t buf_bytes, int exact_start,
                        uint8_t prev0, uint32_t warm_bits, int stream_end, const mh_dectable* dt, uint8_t* d_out,
                        uint64_t out_capacity, uint64_t* d_result, mh_workspace* ws, mh_stream_t stream) {
	return launch_decode_shard(d_bits, start_bit, n_bits, buf_bytes, exact_start, prev0, warm_bits, stream_end, dt, d_out,
	                           out_capacity, reinterpret_cast<unsigned long long*>(d_result), ws, static_cast<cudaStream_t>(stream), 2);
}

uint32_t mh_decode_subsequence_bits(int order, uint64_t n_bits) { return decode_sub_bits(order, n_bits); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// host-buffer session
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kMaxPipeChunks = 4096;   // seam bytes kept per pipelined compress

struct mh_session {
	int device = 0;
	cudaStream_t stream = nullptr;
	uint64_t max_input = 0;      // capacity of d_raw: fixed at creation, never reallocated
	uint64_t enc_chunk = 0;      // bytes one histogram / encode launch may take (what the workspace was sized for)
	uint64_t payload_cap = 0;
	uint64_t pending_out = 0;    // decoded bytes waiting in d_raw for mh_session_fetch
	uint64_t last_count = 0;     // bytes the last mh_session_decompress produced (resident or not)
	uint8_t* d_raw = nullptr;       // uncompressed side
	uint8_t* d_payload = nullptr;   // compressed side (no header byte)
	uint64_t* d_counts = nullptr;   // [65536]
	uint64_t* d_result = nullptr;   // [4]
	uint64_t* h_counts = nullptr;   // pinned [65536]
	uint64_t* h_result = nullptr;   // pinned [4]
	uint8_t* h_seam = nullptr;      // pinned [kMaxPipeChunks]: the first payload byte of every pipelined encode chunk
	mh_workspace* ws = nullptr;
	mh_codebook book;
	mh_dectable dec;
	// pipelined paths: copy streams and their events (created on first use)
	cudaStream_t h2d = nullptr, d2h = nullptr;
	cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

static int ensure_pipe_streams(mh_session* s) {
	if(s->h2d) return MH_OK;
	MH_CUDA(cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking));
	MH_CUDA(cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking));
	for(int i = 0; i < 2; ++i) {
		MH_CUDA(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
		MH_CUDA(cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
	}
	MH_CUDA(cudaMallocHost(&s->h_seam, kMaxPipeChunks));
	return MH_OK;
}

static void drain_session(mh_session* s) {
	if(s->h2d) cudaStreamSynchronize(s->h2d);
	if(s->d2h) cudaStreamSynchronize(s->d2h);
	cudaStreamSynchronize(s->stream);
}
// CUDA call inside a pipelined loop: every error exit waits for the copies that are still reading / writing the caller's buffers
#define MH_CUDA_DRAIN(s, call)                                                      \
	do {                                                                            \
		cudaError_t e_ = (call);                                                    \
		if(e_ != cudaSuccess) { drain_session(s); return ::mh::cuda_fail(e_, #call); } \
	} while(0)

static uint64_t tunable_bytes(Tunable t, uint64_t fallback) {
	const long long v = tunable(t);
	return v > 0 ? uint64_t(v) : fallback;
}

extern "C" {

void* mh_pinned_alloc(size_t bytes) {
	void* p = nullptr;
	if(cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

void mh_pinned_free(void* p) {
	if(p) cudaFreeHost(p);
}

int mh_session_create(int device, uint64_t max_input_bytes, mh_session** out) {
	// An optimal prefix code built from the data's own counts never averages more than 8 bits per byte; a foreign
	// -e table can expand: such an input is then encoded in chunks that fit (mh_session_compress_with_table).
	return mh_session_create_sized(device, max_input_bytes, max_input_bytes + (max_input_bytes >> 3) + 4096, out);
}

int mh_session_create_sized(int device, uint64_t max_input_bytes, uint64_t max_stream_bytes, mh_session** out) {
	if(!out) return MH_ERR_INVALID_ARG;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if(e != cudaSuccess || ndev == 0) { cudaGetLastError(); t_last_error = "no usable CUDA device"; return MH_ERR_NO_DEVICE; }
	if(device < 0 || device >= ndev) return MH_ERR_INVALID_ARG;
	MH_CUDA(cudaSetDevice(device));
	mh_session* s = new(std::nothrow) mh_session;
	if(!s) return MH_ERR_INVALID_ARG;
	s->device = device;
	s->max_input = max_input_bytes;
	s->enc_chunk = max_input_bytes;
	s->payload_cap = ((max_stream_bytes + 64 + 15) / 16) * 16;
	auto fail = [&](int rc) { mh_session_destroy(s); return rc; };
#define S_CUDA(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) return fail(cuda_fail(e_, #call)); } while(0)
	S_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	S_CUDA(cudaMalloc(&s->d_raw, max_input_bytes + 64));
	S_CUDA(cudaMalloc(&s->d_payload, s->payload_cap));
	S_CUDA(cudaMalloc(&s->d_counts, 65536 * sizeof(uint64_t)));
	S_CUDA(cudaMalloc(&s->d_result, 4 * sizeof(uint64_t)));
	S_CUDA(cudaMallocHost(&s->h_counts, 65536 * sizeof(uint64_t)));
	S_CUDA(cudaMallocHost(&s->h_result, 4 * sizeof(uint64_t)));
#undef S_CUDA
	int rc = mh_workspace_create(max_input_bytes, s->payload_cap, &s->ws);
	if(rc != MH_OK) return fail(rc);
	*out = s;
	return MH_OK;
}

void mh_session_destroy(mh_session* s) {
	if(!s) return;
	cudaSetDevice(s->device);
	drain_session(s);
	if(s->d_raw) cudaFree(s->d_raw);
	if(s->d_payload) cudaFree(s->d_payload);
	if(s->d_counts) cudaFree(s->d_counts);
	if(s->d_result) cudaFree(s->d_result);
	if(s->h_counts) cudaFreeHost(s->h_counts);
	if(s->h_result) cudaFreeHost(s->h_result);
	if(s->h_seam) cudaFreeHost(s->h_seam);
	release_book(&s->book);
	release_dec(&s->dec);
	mh_workspace_destroy(s->ws);
	if(s->stream) cudaStreamDestroy(s->stream);
	if(s->h2d) cudaStreamDestroy(s->h2d);
	if(s->d2h) cudaStreamDestroy(s->d2h);
	for(int i = 0; i < 2; ++i) {
		if(s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
		if(s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
	}
	delete s;
}

static int session_histogram_chunked(mh_session* s, const uint8_t* in, uint64_t n, int order, std::vector<uint64_t>& total);

int mh_session_histogram(mh_session* s, const uint8_t* in, uint64_t n, int order, uint64_t* counts) {
	if(!s || !counts || (!in && n) || (order != 0 && order != 1)) return MH_ERR_INVALID_ARG;
	s->pending_out = 0;
	MH_CUDA(cudaSetDevice(s->device));
	if(n > s->enc_chunk || n > s->max_input) {   // larger than the device buffer: the counts add up over chunks
		std::vector<uint64_t> total;
		int rc = session_histogram_chunked(s, in, n, order, total);
		if(rc == MH_OK) memcpy(counts, total.data(), total.size() * sizeof(uint64_t));
		return rc;
	}
	if(n) MH_CUDA(cudaMemcpyAsync(s->d_raw, in, n, cudaMemcpyHostToDevice, s->stream));
	int rc = launch_histogram(s->d_raw, n, MH_PREV0, order, reinterpret_cast<unsigned long long*>(s->d_counts), s->ws, s->stream);
	if(rc != MH_OK) return rc;
	const size_t bins = order ? 65536 : 256;
	MH_CUDA(cudaMemcpyAsync(s->h_counts, s->d_counts, bins * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
	MH_CUDA(cudaStreamSynchronize(s->stream));
	memcpy(counts, s->h_counts, bins * sizeof(uint64_t));
	return MH_OK;
}

// header: 0 0 1 1 E R R R, E = inverse of the coder type, RRR = unused bits of the last byte (src/coding.cpp:88)
static uint8_t stream_header(int order, uint64_t bits) { return uint8_t(0x30 | ((~order & 1) << 3) | ((8 - bits % 8) % 8)); }

// encode d_raw[0..n) with `t`, fetch header + payload into out
static int session_encode(mh_session* s, const mh_table* t, uint64_t n, uint8_t* out, uint64_t out_capacity,
                          uint64_t* out_len, uint64_t* dropped) {
	if(out_capacity < 1) return MH_ERR_CAPACITY;
	int rc = upload_codebook(t, &s->book, s->stream);
	if(rc != MH_OK) return rc;
	rc = launch_encode(s->d_raw, n, MH_PREV0, &s->book, 0, s->d_payload, s->payload_cap,
	                   reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream);
	if(rc != MH_OK) return rc;
	MH_CUDA(cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
	MH_CUDA(cudaStreamSynchronize(s->stream));
	if(s->h_result[2]) return MH_ERR_CAPACITY;
	const uint64_t bits = s->h_result[0];
	const uint64_t bytes = (bits + 7) / 8;
	if(dropped) *dropped = s->h_result[1];
	*out_len = 1 + bytes;
	if(1 + bytes > out_capacity) return MH_ERR_CAPACITY;
	out[0] = stream_header(t->impl.order, bits);
	if(bytes) {
		MH_CUDA(cudaMemcpyAsync(out + 1, s->d_payload, bytes, cudaMemcpyDeviceToHost, s->stream));
		MH_CUDA(cudaStreamSynchronize(s->stream));
	}
	return MH_OK;
}

// ---- inputs larger than the session's device buffer (streaming with carried state) -----------------------
// The input goes through the device buffer in chunks. The histogram adds up over the chunks, each seeded with the
// byte before it; then every chunk is encoded at its global bit offset (bit_base) and its payload lands in the host
// stream at the byte it shares with its neighbour, OR-merged — the same arithmetic that shards one stream over
// several GPUs, applied in sequence on one.
static int session_histogram_chunked(mh_session* s, const uint8_t* in, uint64_t n, int order, std::vector<uint64_t>& total) {
	const size_t bins = order ? 65536 : 256;
	total.assign(bins, 0);
	if(s->enc_chunk == 0) return MH_ERR_CAPACITY;
	for(uint64_t off = 0; off < n; off += s->enc_chunk) {
		const uint64_t len = n - off < s->enc_chunk ? n - off : s->enc_chunk;
		MH_CUDA(cudaMemcpyAsync(s->d_raw, in + off, len, cudaMemcpyHostToDevice, s->stream));
		int rc = launch_histogram(s->d_raw, len, off ? in[off - 1] : uint8_t(MH_PREV0), order, reinterpret_cast<unsigned long long*>(s->d_counts), s->ws, s->stream);
		if(rc != MH_OK) return rc;
		MH_CUDA(cudaMemcpyAsync(s->h_counts, s->d_counts, bins * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
		MH_CUDA(cudaStreamSynchronize(s->stream));
		for(size_t i = 0; i < bins; ++i) total[i] += s->h_counts[i];
	}
	return MH_OK;
}

// `chunk`: input bytes per launch (<= enc_chunk). A table whose codewords expand the input (a foreign -e table) gets
// chunks small enough that chunk x longest codeword fits the compressed-side buffer.
static int session_encode_chunked(mh_session* s, const mh_table* t, const uint8_t* in, uint64_t n, uint64_t chunk, uint8_t* out,
                                  uint64_t out_capacity, uint64_t* out_len, uint64_t* dropped) {
	if(out_capacity < 1 || chunk == 0) return MH_ERR_CAPACITY;
	int rc = upload_codebook(t, &s->book, s->stream);
	if(rc != MH_OK) return rc;
	uint64_t bit_base = 0, drop = 0;
	bool fits = true;
	for(uint64_t off = 0; off < n; off += chunk) {
		const uint64_t len = n - off < chunk ? n - off : chunk;
		MH_CUDA(cudaMemcpyAsync(s->d_raw, in + off, len, cudaMemcpyHostToDevice, s->stream));
		rc = launch_encode(s->d_raw, len, off ? in[off - 1] : uint8_t(MH_PREV0), &s->book, bit_base, s->d_payload, s->payload_cap,
		                   reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream);
		if(rc != MH_OK) return rc;
		MH_CUDA(cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
		MH_CUDA(cudaStreamSynchronize(s->stream));
		if(s->h_result[2]) return MH_ERR_WORKSPACE;   // the device-side buffer: the chunk was sized wrongly
		const uint64_t bits = s->h_result[0];
		drop += s->h_result[1];
		const uint32_t phase = uint32_t(bit_base & 7);
		const uint64_t nbytes = (phase + bits + 7) / 8, first = 1 + (bit_base >> 3);
		if(first + nbytes > out_capacity) fits = false;   // keep counting: the caller learns the size it needs
		if(nbytes && fits) {
			uint8_t seam = 0;
			if(phase) {   // the chunk's first byte is the previous chunk's last: keep that one's bits, add ours
				MH_CUDA(cudaMemcpyAsync(&seam, s->d_payload, 1, cudaMemcpyDeviceToHost, s->stream));
				if(nbytes > 1) MH_CUDA(cudaMemcpyAsync(out + first + 1, s->d_payload + 1, nbytes - 1, cudaMemcpyDeviceToHost, s->stream));
			} else {
				MH_CUDA(cudaMemcpyAsync(out + first, s->d_payload, nbytes, cudaMemcpyDeviceToHost, s->stream));
			}
			MH_CUDA(cudaStreamSynchronize(s->stream));
			if(phase) out[first] = uint8_t(out[first] | (seam & (0xFFu >> phase)));
		}
		bit_base += bits;
	}
	if(dropped) *dropped = drop;
	*out_len = 1 + (bit_base + 7) / 8;
	if(!fits) return MH_ERR_CAPACITY;
	out[0] = stream_header(t->impl.order, bit_base);
	return MH_OK;
}

// ---- pipelined compress (overlapping PCIe with the kernels) -----------------------------------
// Phase 1 (needs the table, so only when `t` is null): the input travels to the device in chunks on the copy stream
// while the histogram of the chunk before it runs — the counts accumulate in one device table, every chunk seeded
// with the byte before it. Phase 2: the resident input is encoded in chunks at their global bit offsets into the two
// halves of the compressed-side buffer; while chunk k is encoded, the payload of chunk k - 1 travels to the host
// (a byte that two chunks share is OR-merged at the end from the chunks' first bytes). With a given table (`t`)
// there is no phase 1: chunk k + 1 travels to the device while chunk k is encoded and chunk k - 1 travels back.
// Returns MH_ERR_WORKSPACE when the buffers are too small for it (the caller takes the sequential path).
static int session_compress_pipelined(mh_session* s, const mh_table* given, const uint8_t* in, uint64_t n, int order, uint8_t* out,
                                      uint64_t out_capacity, uint64_t* out_len, uint64_t* dropped, mh_table** table_out) {
	const uint64_t half = (s->payload_cap / 2) & ~uint64_t(15);
	uint64_t chunk = tunable_bytes(kTunEncPipeChunkBytes, 64ull << 20);
	const int maxb = given ? (given->impl.max_code_bits() > 8 ? given->impl.max_code_bits() : 9) : 9;
	while(chunk > 4096 && (chunk * uint64_t(maxb) + 7) / 8 + 4096 > half) chunk >>= 1;
	chunk &= ~uint64_t(15);
	if(chunk < 4096 || n > s->max_input || n > s->enc_chunk || (n + chunk - 1) / chunk > kMaxPipeChunks || out_capacity < 1) return MH_ERR_WORKSPACE;
	int rc = ensure_pipe_streams(s);
	if(rc != MH_OK) return rc;
	mh_table* built = nullptr;
	const mh_table* t = given;
	if(!given) {
		const size_t bins = order ? 65536 : 256;
		MH_CUDA(cudaMemsetAsync(s->d_counts, 0, bins * sizeof(uint64_t), s->stream));
		for(uint64_t off = 0; off < n; off += chunk) {
			const uint64_t len = n - off < chunk ? n - off : chunk;
			MH_CUDA_DRAIN(s, cudaMemcpyAsync(s->d_raw + off, in + off, len, cudaMemcpyHostToDevice, s->h2d));
			MH_CUDA_DRAIN(s, cudaEventRecord(s->ev_in[0], s->h2d));
			MH_CUDA_DRAIN(s, cudaStreamWaitEvent(s->stream, s->ev_in[0], 0));
			rc = launch_histogram(s->d_raw + off, len, off ? in[off - 1] : uint8_t(MH_PREV0), order,
			                      reinterpret_cast<unsigned long long*>(s->d_counts), s->ws, s->stream, /*accumulate=*/true);
			if(rc != MH_OK) { drain_session(s); return rc; }
		}
		// The encoder's tables are built on the device from the accumulated counts (mh_tables.cu): the first encode chunk
		// follows the last histogram chunk without a host round trip. The counts travel to the host on the D2H stream
		// meanwhile; the host builds its table (table file, header) while that chunk is encoded.
		MH_CUDA_DRAIN(s, cudaEventRecord(s->ev_in[1], s->stream));
		MH_CUDA_DRAIN(s, cudaStreamWaitEvent(s->d2h, s->ev_in[1], 0));
		MH_CUDA_DRAIN(s, cudaMemcpyAsync(s->h_counts, s->d_counts, bins * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->d2h));
		rc = launch_build_codebook(reinterpret_cast<const unsigned long long*>(s->d_counts), order, &s->book, s->stream);
		if(rc != MH_OK) { drain_session(s); return rc; }
	}
	auto fail = [&](int code) { drain_session(s); if(built) mh_table_destroy(built); return code; };
	if(given) {
		rc = upload_codebook(t, &s->book, s->stream);
		if(rc != MH_OK) return fail(rc);
	}
	if(given && n) {   // the first chunk starts its trip now
		const uint64_t len = n < chunk ? n : chunk;
		if(cudaMemcpyAsync(s->d_raw, in, len, cudaMemcpyHostToDevice, s->h2d) != cudaSuccess || cudaEventRecord(s->ev_in[0], s->h2d) != cudaSuccess)
			return fail(cuda_fail(cudaGetLastError(), "pipelined compress: H2D"));
	}
	struct Seam { uint64_t at; uint32_t k, phase; };
	std::vector<Seam> seams;
	uint64_t bit_base = 0, drop = 0;
	uint32_t k = 0;
	for(uint64_t off = 0; off < n; off += chunk, ++k) {
		const uint32_t cur = k & 1;
		const uint64_t len = n - off < chunk ? n - off : chunk;
		if(given) {
			const uint64_t noff = off + chunk;
			if(noff < n) {   // chunk k + 1 travels while chunk k is encoded
				const uint64_t nlen = n - noff < chunk ? n - noff : chunk;
				if(cudaMemcpyAsync(s->d_raw + noff, in + noff, nlen, cudaMemcpyHostToDevice, s->h2d) != cudaSuccess ||
				   cudaEventRecord(s->ev_in[cur ^ 1], s->h2d) != cudaSuccess)
					return fail(cuda_fail(cudaGetLastError(), "pipelined compress: H2D"));
			}
			if(cudaStreamWaitEvent(s->stream, s->ev_in[cur], 0) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "pipelined compress: wait"));
		}
		if(k >= 2 && cudaStreamWaitEvent(s->stream, s->ev_out[cur], 0) != cudaSuccess)   // the payload of chunk k - 2 has left this half
			return fail(cuda_fail(cudaGetLastError(), "pipelined compress: wait"));
		uint8_t* d_half = s->d_payload + cur * half;
		rc = launch_encode(s->d_raw + off, len, off ? in[off - 1] : uint8_t(MH_PREV0), &s->book, bit_base, d_half, half,
		                   reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream);
		if(rc != MH_OK) return fail(rc);
		if(cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess)
			return fail(cuda_fail(cudaGetLastError(), "pipelined compress: result"));
		if(!t) {   // the host's table, built while the first chunk is encoded from the device-built tables
			if(cudaStreamSynchronize(s->d2h) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "pipelined compress: counts"));
			rc = mh_table_from_counts(s->h_counts, order, &built);
			if(rc != MH_OK) return fail(rc);
			t = built;
		}
		if(cudaStreamSynchronize(s->stream) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "pipelined compress: result"));
		if(s->h_result[3]) {   // the device-built tables did not fit the encoder's launch: the host's tables, this chunk again
			rc = upload_codebook(t, &s->book, s->stream);
			if(rc == MH_OK) rc = launch_encode(s->d_raw + off, len, off ? in[off - 1] : uint8_t(MH_PREV0), &s->book, bit_base, d_half, half,
			                                   reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream);
			if(rc != MH_OK) return fail(rc);
			if(cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
			   cudaStreamSynchronize(s->stream) != cudaSuccess)
				return fail(cuda_fail(cudaGetLastError(), "pipelined compress: result"));
		}
		if(s->h_result[2]) return fail(MH_ERR_WORKSPACE);
		const uint64_t bits = s->h_result[0];
		drop += s->h_result[1];
		const uint32_t phase = uint32_t(bit_base & 7);
		const uint64_t nbytes = (phase + bits + 7) / 8, first = 1 + (bit_base >> 3);
		if(first + nbytes > out_capacity) return fail(MH_ERR_WORKSPACE);   // the sequential path reports the size needed
		cudaError_t ce = cudaSuccess;
		if(nbytes) {
			if(phase) {   // the chunk's first byte is the previous chunk's last one: merged on the host when everything has arrived
				ce = cudaMemcpyAsync(s->h_seam + k, d_half, 1, cudaMemcpyDeviceToHost, s->d2h);
				if(ce == cudaSuccess && nbytes > 1) ce = cudaMemcpyAsync(out + first + 1, d_half + 1, nbytes - 1, cudaMemcpyDeviceToHost, s->d2h);
				seams.push_back({first, k, phase});
			} else {
				ce = cudaMemcpyAsync(out + first, d_half, nbytes, cudaMemcpyDeviceToHost, s->d2h);
			}
		}
		if(ce == cudaSuccess) ce = cudaEventRecord(s->ev_out[cur], s->d2h);
		if(ce != cudaSuccess) return fail(cuda_fail(ce, "pipelined compress: D2H"));
		bit_base += bits;
	}
	if(!t) {   // an empty input: no chunk was encoded
		if(cudaStreamSynchronize(s->d2h) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "pipelined compress: counts"));
		rc = mh_table_from_counts(s->h_counts, order, &built);
		if(rc != MH_OK) return fail(rc);
		t = built;
	}
	drain_session(s);
	for(const Seam& sm : seams) out[sm.at] = uint8_t(out[sm.at] | (s->h_seam[sm.k] & (0xFFu >> sm.phase)));
	if(dropped) *dropped = drop;
	*out_len = 1 + (bit_base + 7) / 8;
	out[0] = stream_header(t->impl.order, bit_base);
	if(built) {
		if(table_out) *table_out = built;
		else mh_table_destroy(built);
	}
	return MH_OK;
}

int mh_session_compress(mh_session* s, const uint8_t* in, uint64_t n, int order, uint8_t* out, uint64_t out_capacity,
                        uint64_t* out_len, mh_table** table_out) {
	if(!s || !out || !out_len || (!in && n) || (order != 0 && order != 1)) return MH_ERR_INVALID_ARG;
	s->pending_out = s->last_count = 0;
	MH_CUDA(cudaSetDevice(s->device));
	if(n > s->enc_chunk || n > s->max_input) {   // larger than the device buffer: stream it through in chunks
		std::vector<uint64_t> total;
		int rc = session_histogram_chunked(s, in, n, order, total);
		if(rc != MH_OK) return rc;
		mh_table* t = nullptr;
		rc = mh_table_from_counts(total.data(), order, &t);
		if(rc != MH_OK) return rc;
		rc = session_encode_chunked(s, t, in, n, s->enc_chunk, out, out_capacity, out_len, nullptr);
		if(rc == MH_OK && table_out) *table_out = t;
		else mh_table_destroy(t);
		return rc;
	}
	if(n >= tunable_bytes(kTunPipeMinBytes, 32ull << 20)) {   // large: overlap the copies with the kernels
		const int prc = session_compress_pipelined(s, nullptr, in, n, order, out, out_capacity, out_len, nullptr, table_out);
		if(prc != MH_ERR_WORKSPACE) return prc;
	}
	if(out_capacity < 1) return MH_ERR_CAPACITY;
	int rc = ensure_pipe_streams(s);
	if(rc != MH_OK) return rc;
	if(n) MH_CUDA(cudaMemcpyAsync(s->d_raw, in, n, cudaMemcpyHostToDevice, s->stream));
	rc = launch_histogram(s->d_raw, n, MH_PREV0, order, reinterpret_cast<unsigned long long*>(s->d_counts), s->ws, s->stream);
	if(rc != MH_OK) return rc;
	const size_t bins = order ? 65536 : 256;
	// histogram -> encoder tables on the device -> encoder, without a host round trip; the counts travel to the host on
	// the side stream, where the host's table (table file, header) is built while the encoder runs
	MH_CUDA(cudaEventRecord(s->ev_in[1], s->stream));
	MH_CUDA(cudaStreamWaitEvent(s->d2h, s->ev_in[1], 0));
	MH_CUDA_DRAIN(s, cudaMemcpyAsync(s->h_counts, s->d_counts, bins * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->d2h));
	rc = launch_build_codebook(reinterpret_cast<const unsigned long long*>(s->d_counts), order, &s->book, s->stream);
	if(rc == MH_OK) rc = launch_encode(s->d_raw, n, MH_PREV0, &s->book, 0, s->d_payload, s->payload_cap, reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream);
	if(rc != MH_OK) { drain_session(s); return rc; }
	MH_CUDA_DRAIN(s, cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
	MH_CUDA_DRAIN(s, cudaStreamSynchronize(s->d2h));
	mh_table* t = nullptr;
	rc = mh_table_from_counts(s->h_counts, order, &t);
	if(rc != MH_OK) { drain_session(s); return rc; }
	MH_CUDA_DRAIN(s, cudaStreamSynchronize(s->stream));
	if(s->h_result[3]) {   // the device-built tables did not fit the encoder's launch: the host's tables
		rc = session_encode(s, t, n, out, out_capacity, out_len, nullptr);
	} else if(s->h_result[2]) {
		rc = MH_ERR_CAPACITY;
	} else {
		const uint64_t bits = s->h_result[0], bytes = (bits + 7) / 8;
		*out_len = 1 + bytes;
		if(1 + bytes > out_capacity) rc = MH_ERR_CAPACITY;
		else {
			out[0] = stream_header(order, bits);
			if(bytes) {
				MH_CUDA_DRAIN(s, cudaMemcpyAsync(out + 1, s->d_payload, bytes, cudaMemcpyDeviceToHost, s->stream));
				MH_CUDA_DRAIN(s, cudaStreamSynchronize(s->stream));
			}
		}
	}
	if(rc == MH_OK && table_out) *table_out = t;
	else mh_table_destroy(t);
	return rc;
}

int mh_session_compress_with_table(mh_session* s, const mh_table* t, const uint8_t* in, uint64_t n, uint8_t* out,
                                   uint64_t out_capacity, uint64_t* out_len, uint64_t* dropped) {
	if(!s || !t || !out || !out_len || (!in && n)) return MH_ERR_INVALID_ARG;
	s->pending_out = s->last_count = 0;
	MH_CUDA(cudaSetDevice(s->device));
	// A foreign table may expand the input (the reference just writes the larger file): the chunk is bounded so that
	// chunk x longest codeword fits the compressed-side buffer, and the chunks are coded at their global bit offsets.
	const uint64_t maxb = uint64_t(t->impl.max_code_bits() > 0 ? t->impl.max_code_bits() : 1);
	uint64_t chunk = s->enc_chunk < s->max_input ? s->enc_chunk : s->max_input;
	const uint64_t fit = (s->payload_cap - 64) * 8 / maxb;
	if(chunk > fit) chunk = fit & ~uint64_t(15);
	if(n > chunk) return session_encode_chunked(s, t, in, n, chunk, out, out_capacity, out_len, dropped);
	if(n >= tunable_bytes(kTunPipeMinBytes, 32ull << 20)) {
		const int prc = session_compress_pipelined(s, t, in, n, t->impl.order, out, out_capacity, out_len, dropped, nullptr);
		if(prc != MH_ERR_WORKSPACE) return prc;
	}
	if(n) MH_CUDA(cudaMemcpyAsync(s->d_raw, in, n, cudaMemcpyHostToDevice, s->stream));
	return session_encode(s, t, n, out, out_capacity, out_len, dropped);
}

// A payload larger than the session's device buffer — or one that decodes to more bytes than the uncompressed-side
// buffer holds — is decoded in chunks of bit ranges: every chunk starts from the exact state (bit position, previous
// symbol) its predecessor ended in — the decoder's shard interface with a known start — and keeps 64 bytes of the
// stream behind its range for the codeword that straddles the end. A chunk whose symbols do not fit the buffer is
// halved and decoded again (the session never allocates after creation). The decoded bytes leave for the host chunk
// by chunk; without `out` the pass only counts (nothing stays resident to fetch).
static int session_decompress_chunked(mh_session* s, const mh_table* t, const uint8_t* payload, uint64_t payload_bytes, uint64_t n_bits,
                                      uint8_t* out, uint64_t out_capacity, uint64_t* out_len) {
	const uint64_t cap = s->payload_cap & ~uint64_t(3);
	if(cap < 4096 || s->max_input < 64) return MH_ERR_CAPACITY;
	int rc = upload_dectable(t, &s->dec, s->stream);
	if(rc != MH_OK) return rc;
	uint64_t p = 0, produced = 0;
	uint64_t limit_bits = ~uint64_t(0);   // shrinks when a chunk's symbols outgrow the uncompressed-side buffer
	uint8_t ctx = MH_PREV0;
	int corrupt = 0;
	bool fits = true;
	while(p < n_bits) {
		const uint64_t b0 = (p >> 3) & ~uint64_t(3);
		const uint32_t start_bit = uint32_t(p - 8 * b0);
		const uint64_t load = payload_bytes - b0 < cap ? payload_bytes - b0 : cap;
		bool last = b0 + load >= payload_bytes;
		uint64_t nb = last ? n_bits - p : (load - 64) * 8 - start_bit;
		MH_CUDA(cudaMemcpyAsync(s->d_payload, payload + b0, load, cudaMemcpyHostToDevice, s->stream));
		int iters = 2;
		for(;;) {
			if(nb > limit_bits) { nb = limit_bits; last = false; }
			rc = launch_decode_shard(s->d_payload, start_bit, nb, load, 1, ctx, 0, last ? 1 : 0, &s->dec, s->d_raw, s->max_input,
			                         reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream, iters);
			if(rc != MH_OK) return rc;
			MH_CUDA(cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
			MH_CUDA(cudaStreamSynchronize(s->stream));
			const int64_t dev_status = int64_t(s->h_result[1]);
			if(dev_status == MH_ERR_NOT_CONVERGED && iters < (1 << 20)) { iters *= 4; continue; }
			if(dev_status == MH_ERR_CAPACITY) {   // more symbols than d_raw holds: a smaller bit range, decoded again
				const uint64_t need = s->h_result[0];
				uint64_t shrunk = need ? nb / (need / s->max_input + 1) : nb / 2;   // aim below the buffer size, at least halve
				if(shrunk > nb / 2) shrunk = nb / 2;
				shrunk &= ~uint64_t(63);
				if(shrunk < 512) return MH_ERR_CAPACITY;   // the buffer cannot even hold the symbols of 512 bits
				limit_bits = shrunk;
				continue;
			}
			if(dev_status != 0) return int(dev_status);
			break;
		}
		const uint64_t count = s->h_result[0];
		if(int64_t(s->h_result[2]) != 0) corrupt = int(int64_t(s->h_result[2]));
		if(out && fits) {
			if(produced + count > out_capacity) fits = false;
			else if(count) {
				MH_CUDA(cudaMemcpyAsync(out + produced, s->d_raw, count, cudaMemcpyDeviceToHost, s->stream));
				MH_CUDA(cudaStreamSynchronize(s->stream));
			}
		}
		produced += count;
		if(last) break;
		const uint32_t end = uint32_t(s->h_result[3] & 0xffffffffull);   // where the chunk's last codeword ended, and that symbol
		p += nb + (end >> 8);
		ctx = uint8_t(end & 255u);
	}
	*out_len = produced;
	s->pending_out = 0;
	s->last_count = produced;
	if(out && !fits) return MH_ERR_CAPACITY;
	return corrupt;
}

// Extract straight into a host buffer, pipelined: the payload is cut into bit-range chunks (as above); while chunk k is
// decoded, chunk k + 1 travels to the device and the bytes of chunk k - 1 travel back — PCIe is full duplex, so the
// 2.4x larger D2H side hides both the H2D copies and the kernels. Two halves of the payload buffer and of the
// uncompressed-side buffer alternate. Returns MH_ERR_WORKSPACE when the buffers are too small for it (the caller then
// takes the sequential path); needs pinned host memory to actually overlap.
static int session_decompress_pipelined(mh_session* s, const mh_table* t, const uint8_t* payload, uint64_t payload_bytes, uint64_t n_bits,
                                        uint8_t* out, uint64_t out_capacity, uint64_t* out_len) {
	const uint64_t pay_half = (s->payload_cap / 2) & ~uint64_t(15), out_half = (s->max_input / 2) & ~uint64_t(63);
	uint64_t chunk = tunable_bytes(kTunPipeChunkBytes, 16ull << 20) & ~uint64_t(3);   // measured best of 16, 32, 64, 128 MiB
	if(chunk > pay_half) chunk = pay_half & ~uint64_t(3);
	if(chunk < 4096 || out_half < 4096) return MH_ERR_WORKSPACE;
	int rc = ensure_pipe_streams(s);
	if(rc != MH_OK) return rc;
	auto geometry = [&](uint64_t base_bit, uint64_t& b0, uint64_t& load, bool& last) {   // chunk that starts (nominally) at base_bit
		b0 = (base_bit >> 3) & ~uint64_t(3);
		load = payload_bytes - b0 < chunk ? payload_bytes - b0 : chunk;
		last = b0 + load >= payload_bytes;
	};
	uint64_t base = 0, rel = 0, produced = 0, b0, load;
	uint8_t ctx = MH_PREV0;
	bool last;
	int corrupt = 0;
	geometry(0, b0, load, last);
	MH_CUDA_DRAIN(s, cudaMemcpyAsync(s->d_payload, payload + b0, load, cudaMemcpyHostToDevice, s->h2d));
	MH_CUDA_DRAIN(s, cudaEventRecord(s->ev_in[0], s->h2d));
	rc = upload_dectable(t, &s->dec, s->stream);   // flattened on the host while the first chunk is on its way
	if(rc != MH_OK) { drain_session(s); return rc; }
	for(uint32_t k = 0;; ++k) {
		const uint32_t cur = k & 1;
		geometry(base, b0, load, last);
		const uint64_t p = base + rel;                      // exact first bit: where the previous chunk's last codeword ended
		const uint64_t first = p - 8 * b0;                  // ... relative to the bytes in the buffer
		const uint64_t skip = (first >> 5) * 4;             // whole words before it
		const uint64_t nb = last ? n_bits - p : (load - 64) * 8 - first;
		const uint64_t next_base = 8 * (b0 + load - 64);    // the range ends 64 bytes before the loaded bytes do
		if(!last) {                                         // the next chunk's bytes do not depend on this chunk's result
			uint64_t nb0, nload;
			bool nlast;
			geometry(next_base, nb0, nload, nlast);
			MH_CUDA_DRAIN(s, cudaMemcpyAsync(s->d_payload + (cur ^ 1) * pay_half, payload + nb0, nload, cudaMemcpyHostToDevice, s->h2d));
			MH_CUDA_DRAIN(s, cudaEventRecord(s->ev_in[cur ^ 1], s->h2d));
		}
		MH_CUDA_DRAIN(s, cudaStreamWaitEvent(s->stream, s->ev_in[cur], 0));
		if(k >= 2) MH_CUDA_DRAIN(s, cudaStreamWaitEvent(s->stream, s->ev_out[cur], 0));   // the bytes of chunk k - 2 have left this half
		int iters = 2;
		for(;;) {
			rc = launch_decode_shard(s->d_payload + cur * pay_half + skip, uint32_t(first & 31), nb, load - skip, 1, ctx, 0, last ? 1 : 0, &s->dec,
			                         s->d_raw + cur * out_half, out_half, reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream, iters);
			if(rc != MH_OK) { drain_session(s); return rc; }
			MH_CUDA_DRAIN(s, cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
			MH_CUDA_DRAIN(s, cudaStreamSynchronize(s->stream));
			const int64_t dev_status = int64_t(s->h_result[1]);
			if(dev_status == MH_ERR_NOT_CONVERGED && iters < (1 << 20)) { iters *= 4; continue; }
			if(dev_status == MH_ERR_CAPACITY) { drain_session(s); return MH_ERR_WORKSPACE; }   // a chunk outgrew its half: sequential path
			if(dev_status != 0) { drain_session(s); return int(dev_status); }
			break;
		}
		const uint64_t count = s->h_result[0];
		if(int64_t(s->h_result[2]) != 0) corrupt = int(int64_t(s->h_result[2]));
		if(produced + count > out_capacity) { drain_session(s); return MH_ERR_WORKSPACE; }   // the sequential path reports the size needed
		if(count) MH_CUDA_DRAIN(s, cudaMemcpyAsync(out + produced, s->d_raw + cur * out_half, count, cudaMemcpyDeviceToHost, s->d2h));
		MH_CUDA_DRAIN(s, cudaEventRecord(s->ev_out[cur], s->d2h));
		produced += count;
		if(last) break;
		const uint32_t end = uint32_t(s->h_result[3] & 0xffffffffull);
		rel = end >> 8;
		ctx = uint8_t(end & 255u);
		base = next_base;
	}
	drain_session(s);
	*out_len = produced;
	s->pending_out = 0;
	s->last_count = produced;
	return corrupt;
}

int mh_session_decompress(mh_session* s, const mh_table* t, const uint8_t* stream, uint64_t stream_len, uint8_t* out,
                          uint64_t out_capacity, uint64_t* out_len) {
	if(!s || !t || !stream || !out_len) return MH_ERR_INVALID_ARG;
	if(stream_len < 1) return MH_ERR_BAD_HEADER;
	const uint8_t header = stream[0];
	if((header & 0xF0) != 0x30) return MH_ERR_BAD_HEADER;                       // src/coding.cpp:103-106
	if(((~header >> 3) & 1) != t->impl.order) return MH_ERR_TYPE_MISMATCH;      // src/coding.cpp:107-110
	const uint64_t payload_bytes = stream_len - 1;
	const uint64_t remainder = header & 7;
	// src/coding.cpp:115 in 64-bit; a negative length makes the reference's loop (:124) decode nothing
	const uint64_t n_bits = payload_bytes * 8 < remainder ? 0 : payload_bytes * 8 - remainder;
	MH_CUDA(cudaSetDevice(s->device));
	s->pending_out = s->last_count = 0;
	if(out && payload_bytes >= tunable_bytes(kTunPipeMinBytes, 32ull << 20)) {   // large and wanted on the host: overlap the copies
		const int prc = session_decompress_pipelined(s, t, stream + 1, payload_bytes, n_bits, out, out_capacity, out_len);
		if(prc != MH_ERR_WORKSPACE) return prc;
		s->pending_out = s->last_count = 0;
	}
	if(payload_bytes > s->payload_cap)   // larger than the device buffer: decode it in chunks
		return session_decompress_chunked(s, t, stream + 1, payload_bytes, n_bits, out, out_capacity, out_len);
	int rc = upload_dectable(t, &s->dec, s->stream);
	if(rc != MH_OK) return rc;
	if(payload_bytes) MH_CUDA(cudaMemcpyAsync(s->d_payload, stream + 1, payload_bytes, cudaMemcpyHostToDevice, s->stream));
	int iters = 2;
	for(;;) {
		rc = launch_decode(s->d_payload, 0, n_bits, MH_PREV0, &s->dec, s->d_raw, s->max_input,
		                   reinterpret_cast<unsigned long long*>(s->d_result), s->ws, s->stream, iters);
		if(rc != MH_OK) return rc;
		MH_CUDA(cudaMemcpyAsync(s->h_result, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
		MH_CUDA(cudaStreamSynchronize(s->stream));
		const int64_t dev_status = int64_t(s->h_result[1]);
		if(dev_status == MH_ERR_NOT_CONVERGED && iters < (1 << 20)) { iters *= 4; continue; }   // pathological seams: keep fixing
		break;
	}
	if(int64_t(s->h_result[1]) == MH_ERR_CAPACITY) {
		// The decoded size is only known after the count pass, and it exceeds the uncompressed-side buffer. The session
		// does not reallocate: the stream is decoded in bit-range chunks that fit, and the bytes leave chunk by chunk.
		// Without `out` this only reports the size (mh_session_fetch then answers MH_ERR_WORKSPACE: call again with out).
		if(!out) {
			*out_len = s->last_count = s->h_result[0];
			return MH_OK;
		}
		return session_decompress_chunked(s, t, stream + 1, payload_bytes, n_bits, out, out_capacity, out_len);
	}
	*out_len = s->last_count = s->h_result[0];
	if(int64_t(s->h_result[1]) != 0) return int(int64_t(s->h_result[1]));
	s->pending_out = s->h_result[0];
	const int corrupt = int(int64_t(s->h_result[2]));   // bytes are delivered, like the reference, but flagged
	if(!out) return corrupt;
	rc = mh_session_fetch(s, out, out_capacity, out_len);
	return rc != MH_OK ? rc : corrupt;
}

int mh_session_fetch(mh_session* s, uint8_t* out, uint64_t out_capacity, uint64_t* out_len) {
	if(!s || !out_len) return MH_ERR_INVALID_ARG;
	if(s->pending_out == 0 && s->last_count != 0) {   // decoded in chunks: counted, but nothing stayed on the device
		*out_len = 0;
		return MH_ERR_WORKSPACE;
	}
	if(!out && s->pending_out) return MH_ERR_INVALID_ARG;
	*out_len = s->pending_out;
	if(s->pending_out > out_capacity) return MH_ERR_CAPACITY;
	MH_CUDA(cudaSetDevice(s->device));
	if(s->pending_out) {
		MH_CUDA(cudaMemcpyAsync(out, s->d_raw, s->pending_out, cudaMemcpyDeviceToHost, s->stream));
		MH_CUDA(cudaStreamSynchronize(s->stream));
	}
	return MH_OK;
}

}  // extern "C"
