// mh_shard.cu — ONE logical stream over several GPUs, in the library: mh_comm_* / mh_sharded_*.
//
// The reference has no multi-device path (its callers of compress / decompress are src/main.cpp:204-212); this is what
// a multi-GPU host driver binds in their place. One host thread (or process) per GPU; each owns an mh_comm.
//
// compress (byte-range shards)
//   1. every rank counts its byte range (mh_gpu_histogram's kernels), guessing ' ' as the byte before it, and appends
//      its first and last byte: ONE all-gather of 65,538 x u64 per rank carries everything the phase needs;
//   2. a kernel sums the gathered histograms and moves the one seam pair per shard to the context it really has (the
//      predecessor's last byte), in the sum and in the shard's own row; the encoder's tables are built from the sum ON
//      THE DEVICE (mh_tables.cu), identically on every rank — no host round trip between the gather and the encoder;
//      the sum also travels to the host on a side stream, where the (identical) host table is built while the encoder
//      runs (table file, decoder tables);
//   3. a kernel multiplies every rank's counts with the code lengths of the codebook: every rank knows every
//      shard's payload size and bit offset without another exchange, and the encoder reads its bit phase and its
//      first context from device words (launch_encode's d_bit_base / d_prev0) — the host learns the layout when the
//      encoder has finished.
// decompress (bit-range shards)
//   exact mode: the layout's cuts are codeword boundaries with known contexts (what compress produces): no exchange
//   at all beyond one all-gather of the symbol counts (output offsets).
//   speculative mode (a stream without an index cut at arbitrary bits; what bench.py times): neighbours exchange a
//   halo in ONE all-gather (each shard's last 8192 + 64 bits and first 64 bytes), a kernel splices it around the local
//   payload (OR-merging the byte two shards share), every rank but the first starts 8192 bits before its range from a
//   guessed state (mh_gpu_decode_shard), and ONE all-gather of 32 bytes per rank carries symbol counts and seam
//   states: rank g is accepted when the state its warm-up reached at its first bit equals the state rank g-1 ended in,
//   otherwise it decodes again from exactly that state (never needed so far).
//
// Transport: NCCL (loaded at run time: libnccl.so.2 — NVLink / NVSwitch between the GPUs of a box), or, for ranks that
// live in one process, an in-process transport (device-to-device copies, host barrier) that also lets several ranks
// share one GPU — which is how the single-GPU test suite covers this file.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

#include "mh_host.hpp"
#include "mh_internal.hpp"

namespace mh {
void set_last_error(const char* msg);   // mh_api.cu
int upload_codebook_for(const mh_table* t, mh_codebook* cb, cudaStream_t st);
int upload_dectable_for(const mh_table* t, mh_dectable* dt, cudaStream_t st);
void release_codebook(mh_codebook* cb);
void release_dectable(mh_dectable* dt);
int table_order(const mh_table* t);
uint64_t table_serial(const mh_table* t);   // unique per mh_table object ever created in this process

namespace {

// ---- NCCL, resolved at run time so that libmh_gpu.so loads (and the single-GPU paths work) without it ------------
struct NcclApi {
	bool ok = false;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl_api() {
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		void* h = nullptr;
		for(const char* name : {"libnccl.so.2", "libnccl.so"}) {
			h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if(h) break;
		}
		if(!h) return;
		api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
		api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
		api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
		api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
		api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
		api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
		api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather && api.GetErrorString;
	});
	return api;
}

int nccl_fail(ncclResult_t r, const char* what) {
	std::string msg = std::string(what) + ": " + (nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "NCCL error");
	set_last_error(msg.c_str());
	return MH_ERR_CUDA;
}
#define MH_NCCL(call)                                            \
	do {                                                         \
		ncclResult_t r_ = (call);                                \
		if(r_ != ncclSuccess) return nccl_fail(r_, #call);       \
	} while(0)

// ---- in-process transport ---------------------------------------------------------------------------------------
struct LocalGroup {
	explicit LocalGroup(int w) : world(w), send(w, nullptr) {}
	const int world;
	std::mutex mu;
	std::condition_variable cv;
	int arrived = 0;
	uint64_t generation = 0;
	std::vector<const void*> send;
	// ranks of one process working on shared host buffers (mh_sharded_*_host): seam bytes, per-rank values
	std::vector<uint8_t> seam = std::vector<uint8_t>(MH_MAX_SHARDS, 0);
	std::vector<uint64_t> value = std::vector<uint64_t>(MH_MAX_SHARDS, 0);
	std::vector<int> status = std::vector<int>(MH_MAX_SHARDS, 0);
	void barrier() {
		std::unique_lock<std::mutex> lock(mu);
		const uint64_t gen = generation;
		if(++arrived == world) {
			arrived = 0;
			++generation;
			cv.notify_all();
		} else {
			cv.wait(lock, [&] { return generation != gen; });
		}
	}
};

constexpr uint32_t kMsgWords = 65536 + 2;        // histogram, first byte, last byte (0x100: the shard is empty)
constexpr uint32_t kShardPad = 4096;             // bytes in front of a shard's payload inside its local buffer (room for the warm-up halo)
constexpr uint32_t kHaloHead = 64;               // bytes a shard lends its predecessor so that one's last codeword can complete
constexpr uint32_t kWarmBits = MH_DECODE_WARM_UNIT;
constexpr uint32_t kHaloTail = kWarmBits / 8 + 8;   // bytes a shard lends its successor: the warm-up, plus alignment slack
static_assert(kHaloTail + 8 <= kShardPad, "the warm-up halo must fit in front of the payload");

// ---- kernels -------------------------------------------------------------------------------------------------
__global__ void shard_edges_kernel(const uint8_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ edges) {
	edges[0] = n ? in[0] : 0x100ull;
	edges[1] = n ? in[n - 1] : 0x100ull;
}

// total[i] = sum over ranks of gathered[r][i]
__global__ void __launch_bounds__(256) shard_sum_kernel(const unsigned long long* __restrict__ gathered, uint32_t world, uint32_t bins,
                                                         unsigned long long* __restrict__ total) {
	const uint32_t i = blockIdx.x * 256 + threadIdx.x;
	if(i >= bins) return;
	unsigned long long s = 0;
	for(uint32_t r = 0; r < world; ++r) s += gathered[size_t(r) * kMsgWords + i];
	total[i] = s;
}

// Every shard counted its first byte as following ' '; move that one count to the pair it really forms with the last
// byte before it (order 1), in the shard's own row and in the sum. tail[g] = the byte before shard g (its prev0).
__global__ void shard_seam_kernel(unsigned long long* __restrict__ gathered, uint32_t world, int order, unsigned long long* __restrict__ total,
                                  unsigned long long* __restrict__ tail) {
	if(threadIdx.x || blockIdx.x) return;
	uint32_t prev = MH_PREV0;
	for(uint32_t g = 0; g < world; ++g) {
		unsigned long long* row = gathered + size_t(g) * kMsgWords;
		tail[g] = prev;
		const unsigned long long first = row[65536], last = row[65537];
		if(first > 255) continue;   // empty shard: the context passes through
		if(order && prev != MH_PREV0) {
			const uint32_t f = uint32_t(first);
			row[256 * MH_PREV0 + f] -= 1;
			row[256 * prev + f] += 1;
			total[256 * MH_PREV0 + f] -= 1;
			total[256 * prev + f] += 1;
		}
		prev = uint32_t(last);
	}
}

// bits[r] = sum over (prev, c) of gathered[r][prev, c] x code length (from the wide codebook: len << 56 | code)
__global__ void __launch_bounds__(256) shard_bits_kernel(const unsigned long long* __restrict__ gathered, const unsigned long long* __restrict__ enc,
                                                          uint32_t bins, unsigned long long* __restrict__ bits) {
	__shared__ unsigned long long part[8];
	const unsigned long long* row = gathered + size_t(blockIdx.x) * kMsgWords;
	unsigned long long s = 0;
	for(uint32_t i = threadIdx.x; i < bins; i += 256) {
		const unsigned long long c = row[i];
		if(c) s += c * (enc[i] >> 56);
	}
	for(int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
	if((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
	__syncthreads();
	if(threadIdx.x == 0) {
		unsigned long long t = 0;
		for(int w = 0; w < 8; ++w) t += part[w];
		bits[blockIdx.x] = t;
	}
}

// layout[0 .. world) = exclusive scan of bits (every shard's global bit offset), layout[world .. 2 world) = bits
__global__ void shard_scan_kernel(const unsigned long long* __restrict__ bits, uint32_t world, unsigned long long* __restrict__ layout) {
	if(threadIdx.x || blockIdx.x) return;
	unsigned long long base = 0;
	for(uint32_t r = 0; r < world; ++r) {
		layout[r] = base;
		layout[world + r] = bits[r];
		base += bits[r];
	}
}

// my_halo = the shard's last kHaloTail payload bytes, then its first kHaloHead
__global__ void halo_pack_kernel(const uint8_t* __restrict__ pay, uint64_t nbytes, uint8_t* __restrict__ my_halo) {
	for(uint32_t i = threadIdx.x; i < kHaloTail + kHaloHead; i += blockDim.x) {
		uint8_t v = 0;
		if(i < kHaloTail) {
			const int64_t at = int64_t(nbytes) - int64_t(kHaloTail) + i;
			if(at >= 0) v = pay[at];
		} else if(i - kHaloTail < nbytes) {
			v = pay[i - kHaloTail];
		}
		my_halo[i] = v;
	}
}

// Splice the neighbours' halos around the local payload: the predecessor's tail in front of it, the successor's head
// behind it; a byte that two shards share (bit phase != 0) is the OR of the two.
__global__ void halo_splice_kernel(uint8_t* __restrict__ local, const uint8_t* __restrict__ halos, uint32_t rank, uint32_t world, uint32_t phase,
                                   uint32_t next_phase, uint64_t nbytes) {
	uint8_t* pay = local + kShardPad;
	const uint32_t per = kHaloTail + kHaloHead;
	if(rank > 0) {
		const uint8_t* tail = halos + size_t(rank - 1) * per;
		for(uint32_t i = threadIdx.x; i < kHaloTail; i += blockDim.x) {
			if(phase) {   // the predecessor's last byte is my first byte
				if(i + 1 < kHaloTail) local[kShardPad - kHaloTail + 1 + i] = tail[i];
				else pay[0] |= tail[i];
			} else {
				local[kShardPad - kHaloTail + i] = tail[i];
			}
		}
	}
	for(uint32_t i = threadIdx.x; i < kHaloHead; i += blockDim.x) {
		uint8_t v = 0;
		if(rank + 1 < world) v = halos[size_t(rank + 1) * per + kHaloTail + i];
		if(rank + 1 < world && next_phase) {
			if(i == 0) pay[nbytes - 1] |= v;
			else pay[nbytes - 1 + i] = v;
		} else {
			pay[nbytes + i] = v;
		}
	}
}

int cuda_fail_if(cudaError_t e, const char* what) { return e == cudaSuccess ? MH_OK : cuda_fail(e, what); }

double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

}  // namespace mh

using namespace mh;

enum { kStatGather = 0, kStatHalo, kStatSeam, kStatTrees, kStatCodebook, kStatDectable, kStatRounds, kStatCalls, kStatCount };

struct mh_comm {
	int device = 0, rank = 0, world = 1;
	ncclComm_t nccl = nullptr;
	std::shared_ptr<LocalGroup> local;   // in-process transport (null with NCCL or a single rank)
	std::shared_ptr<LocalGroup> host;    // ranks created together in one process share this (host barrier, seam bytes)
	// device buffers of the host-buffer calls (mh_sharded_*_host), grown on demand outside the hot path
	cudaStream_t stream = nullptr;
	cudaStream_t side = nullptr;         // carries the summed counts to the host while the caller's stream encodes
	cudaEvent_t ev_side = nullptr;
	uint8_t *d_hin = nullptr, *d_hlocal = nullptr, *d_hout = nullptr;
	uint64_t hin_cap = 0, hlocal_cap = 0, hout_cap = 0;
	// per-rank state, sized on first use (mh_comm_reserve) — nothing is allocated on the hot path afterwards
	mh_workspace* ws = nullptr;
	uint64_t ws_input = 0, ws_payload = 0;
	mh_codebook book;
	mh_dectable dec;
	uint64_t dec_serial = 0;             // serial of the mh_table the decode tables were last flattened from (0: none)
	unsigned long long* d_msg = nullptr;      // [kMsgWords]
	unsigned long long* d_gather = nullptr;   // [world * kMsgWords]
	unsigned long long* d_total = nullptr;    // [65536 + world]: the summed histogram, then every shard's prev0
	unsigned long long* d_bits = nullptr;     // [world]
	unsigned long long* d_layout = nullptr;   // [2 * world]: bit offsets, bit counts
	unsigned long long* d_result = nullptr;   // [4] encode / decode result
	unsigned long long* d_seams = nullptr;    // [4 * world]
	uint8_t* d_halo = nullptr;                // [kHaloTail + kHaloHead]
	uint8_t* d_halos = nullptr;               // [world * (kHaloTail + kHaloHead)]
	unsigned long long* h_total = nullptr;    // pinned [65536 + world]
	unsigned long long* h_small = nullptr;    // pinned [2 * world + 4 + 4 * world]: layout, result, seams
	cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	double stats[kStatCount] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

int comm_alloc(mh_comm* c) {
	const size_t w = size_t(c->world);
	MH_CUDA(cudaSetDevice(c->device));
	MH_CUDA(cudaMalloc(&c->d_msg, kMsgWords * sizeof(unsigned long long)));
	MH_CUDA(cudaMalloc(&c->d_gather, w * kMsgWords * sizeof(unsigned long long)));
	MH_CUDA(cudaMalloc(&c->d_total, (65536 + w) * sizeof(unsigned long long)));
	MH_CUDA(cudaMalloc(&c->d_bits, w * sizeof(unsigned long long)));
	MH_CUDA(cudaMalloc(&c->d_layout, 2 * w * sizeof(unsigned long long)));
	MH_CUDA(cudaMalloc(&c->d_result, 4 * sizeof(unsigned long long)));
	MH_CUDA(cudaMalloc(&c->d_seams, 4 * w * sizeof(unsigned long long)));
	MH_CUDA(cudaMalloc(&c->d_halo, kHaloTail + kHaloHead));
	MH_CUDA(cudaMalloc(&c->d_halos, w * (kHaloTail + kHaloHead)));
	MH_CUDA(cudaMallocHost(&c->h_total, (65536 + w) * sizeof(unsigned long long)));
	MH_CUDA(cudaMallocHost(&c->h_small, (6 * w + 4) * sizeof(unsigned long long)));
	MH_CUDA(cudaMemset(c->d_msg, 0, kMsgWords * sizeof(unsigned long long)));
	for(auto& e : c->ev) MH_CUDA(cudaEventCreate(&e));
	return MH_OK;
}

// all-gather of `bytes` per rank on `st`; the elapsed device time between ev[slot] and ev[slot + 1] is the collective
int all_gather(mh_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t st, int ev_slot) {
	MH_CUDA(cudaEventRecord(c->ev[ev_slot], st));
	if(c->nccl) {
		MH_NCCL(nccl_api().AllGather(send, recv, bytes, ncclUint8, c->nccl, st));
	} else if(c->local) {
		LocalGroup& g = *c->local;
		MH_CUDA(cudaStreamSynchronize(st));   // my contribution is complete
		g.send[c->rank] = send;
		g.barrier();
		for(int r = 0; r < c->world; ++r)
			MH_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(recv) + size_t(r) * bytes, g.send[r], bytes, cudaMemcpyDefault, st));
		MH_CUDA(cudaStreamSynchronize(st));
		g.barrier();                          // nobody rewrites its contribution before everybody has read it
	} else {
		MH_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));   // world == 1
	}
	MH_CUDA(cudaEventRecord(c->ev[ev_slot + 1], st));
	return MH_OK;
}

void add_elapsed(mh_comm* c, int stat, int ev_slot) {
	float ms = 0.f;
	if(cudaEventElapsedTime(&ms, c->ev[ev_slot], c->ev[ev_slot + 1]) == cudaSuccess) c->stats[stat] += double(ms) * 1e3;
	else cudaGetLastError();
}

// Decode the bits [own_bit, own_bit + bits) of this rank's local buffer (bit coordinates of d_local; buf_end readable
// bytes) and agree with the neighbours. spec == false: the start is exact (own_bit is a codeword boundary, prev0 the byte
// before it, and the range ends on a codeword boundary). spec == true: every rank but the first starts kWarmBits before
// own_bit from a guessed state; ONE all-gather of 32 bytes per rank carries symbol counts and seam states; a rank whose
// warm-up did not reach the state its predecessor ended in decodes again from exactly that state. All ranks track what
// every rank started from, so the handshake needs no second message. counts[g] = symbols rank g owns.
int decode_with_handshake(mh_comm* c, uint8_t* d_local, uint64_t own_bit, uint64_t bits, uint64_t buf_end, bool spec, uint8_t first_prev0,
                          uint8_t* d_out, uint64_t out_capacity, cudaStream_t st, bool time_halo, uint64_t* counts) {
	const uint32_t world = uint32_t(c->world), r = uint32_t(c->rank);
	unsigned long long* h_seams = c->h_small + 2 * world + 4;
	bool exact = !spec || r == 0;
	uint8_t prev0 = first_prev0;
	uint64_t start = own_bit;
	uint32_t warm = exact ? 0u : kWarmBits;
	std::vector<int64_t> started(world, -1);
	std::vector<uint64_t> views(world), ends(world);
	int rounds = 0, rc = MH_OK;
	for(;;) {
		const uint64_t origin = start - warm;
		const uint64_t off = (origin / 32) * 4;
		const uint64_t n_bits = own_bit + bits - origin;
		if(bits == 0 || own_bit + bits <= start) {
			MH_CUDA(cudaMemsetAsync(c->d_result, 0, 4 * sizeof(unsigned long long), st));
		} else {
			rc = launch_decode_shard(d_local + off, uint32_t(origin % 32), n_bits, buf_end - off, exact ? 1 : 0, prev0, warm,
			                         (!spec || r == world - 1) ? 1 : 0, &c->dec, d_out, out_capacity, c->d_result, c->ws, st, 2);
			if(rc != MH_OK) return rc;
		}
		rc = all_gather(c, c->d_result, c->d_seams, 4 * sizeof(unsigned long long), st, 4);
		if(rc != MH_OK) return rc;
		MH_CUDA(cudaMemcpyAsync(h_seams, c->d_seams, 4 * world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		MH_CUDA(cudaStreamSynchronize(st));
		if(time_halo) { add_elapsed(c, kStatHalo, 2); time_halo = false; }
		add_elapsed(c, kStatSeam, 4);
		int worst = 0;
		for(uint32_t g = 0; g < world; ++g) {
			const int64_t s1 = int64_t(h_seams[4 * g + 1]), s2 = int64_t(h_seams[4 * g + 2]);
			if(s1 != 0 && (worst == 0 || s1 == MH_ERR_NOT_CONVERGED)) worst = int(s1);
			if(s1 == 0 && s2 != 0 && worst == 0) worst = int(s2);
			counts[g] = h_seams[4 * g];
			views[g] = h_seams[4 * g + 3] >> 32;
			ends[g] = h_seams[4 * g + 3] & 0xffffffffull;
		}
		if(worst != 0) return worst;   // every rank sees the same words and leaves together
		if(!spec) break;
		bool all_ok = true, mine_ok = true;
		std::vector<int64_t> next_started = started;
		for(uint32_t g = 1; g < world; ++g) {
			const bool ok = started[g] >= 0 ? uint64_t(started[g]) == ends[g - 1] : views[g] == ends[g - 1];
			if(!ok) {
				all_ok = false;
				next_started[g] = int64_t(ends[g - 1]);
				if(g == r) mine_ok = false;
			}
		}
		if(all_ok) break;
		if(++rounds > int(world) + 1) { set_last_error("sharded decompress: the seam handshake did not converge"); return MH_ERR_NOT_CONVERGED; }
		started = next_started;
		if(!mine_ok) {   // decode again, this time from the state my predecessor really ended in
			const uint64_t e = ends[r - 1];
			exact = true;
			prev0 = uint8_t(e & 255u);
			warm = 0;
			start = own_bit + (e >> 8);
		}
	}
	c->stats[kStatRounds] += rounds;
	return MH_OK;
}

int ensure_device(uint8_t** p, uint64_t* cap, uint64_t want) {
	if(*cap >= want) return MH_OK;
	if(*p) cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	MH_CUDA(cudaMalloc(p, want));
	*cap = want;
	return MH_OK;
}

}  // namespace

extern "C" {

int mh_comm_available(void) { return nccl_api().ok ? 1 : 0; }

int mh_comm_unique_id(uint8_t id[MH_COMM_ID_BYTES]) {
	if(!id) return MH_ERR_INVALID_ARG;
	if(!nccl_api().ok) { set_last_error("NCCL (libnccl.so.2) could not be loaded"); return MH_ERR_CUDA; }
	static_assert(MH_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
	ncclUniqueId u;
	MH_NCCL(nccl_api().GetUniqueId(&u));
	memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
	return MH_OK;
}

int mh_comm_create(int device, int rank, int world, const uint8_t id[MH_COMM_ID_BYTES], mh_comm** out) {
	if(!out || world < 1 || world > MH_MAX_SHARDS || rank < 0 || rank >= world || (world > 1 && !id)) return MH_ERR_INVALID_ARG;
	int ndev = 0;
	if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_last_error("no usable CUDA device"); return MH_ERR_NO_DEVICE; }
	if(device < 0 || device >= ndev) return MH_ERR_INVALID_ARG;
	if(world > 1 && !nccl_api().ok) { set_last_error("NCCL (libnccl.so.2) could not be loaded"); return MH_ERR_CUDA; }
	mh_comm* c = new(std::nothrow) mh_comm;
	if(!c) return MH_ERR_INVALID_ARG;
	c->device = device; c->rank = rank; c->world = world;
	int rc = comm_alloc(c);
	if(rc == MH_OK && world > 1) {
		ncclUniqueId u;
		memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
		ncclResult_t r = nccl_api().CommInitRank(&c->nccl, world, u, rank);
		if(r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitRank");
	}
	if(rc != MH_OK) { mh_comm_destroy(c); return rc; }
	*out = c;
	return MH_OK;
}

int mh_comm_create_local(int world, const int* devices, int use_nccl, mh_comm** out) {
	if(!out || world < 1 || world > MH_MAX_SHARDS) return MH_ERR_INVALID_ARG;
	int ndev = 0;
	if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_last_error("no usable CUDA device"); return MH_ERR_NO_DEVICE; }
	std::vector<int> dev(world);
	for(int r = 0; r < world; ++r) {
		dev[r] = devices ? devices[r] : r % ndev;
		if(dev[r] < 0 || dev[r] >= ndev) return MH_ERR_INVALID_ARG;
	}
	bool distinct = true;
	for(int a = 0; a < world; ++a)
		for(int b = a + 1; b < world; ++b) distinct = distinct && dev[a] != dev[b];
	const bool with_nccl = use_nccl && world > 1 && distinct && nccl_api().ok;   // NCCL wants one rank per device
	if(use_nccl > 1 && world > 1 && !with_nccl) { set_last_error("NCCL requested but not usable (not loaded, or two ranks share a device)"); return MH_ERR_CUDA; }
	std::shared_ptr<LocalGroup> host_group = std::make_shared<LocalGroup>(world);
	std::shared_ptr<LocalGroup> group = (world > 1 && !with_nccl) ? host_group : nullptr;
	std::vector<ncclComm_t> comms(world, nullptr);
	if(with_nccl) MH_NCCL(nccl_api().CommInitAll(comms.data(), world, dev.data()));
	for(int r = 0; r < world; ++r) out[r] = nullptr;
	for(int r = 0; r < world; ++r) {
		mh_comm* c = new(std::nothrow) mh_comm;
		int rc = c ? MH_OK : MH_ERR_INVALID_ARG;
		if(c) {
			c->device = dev[r]; c->rank = r; c->world = world;
			c->nccl = comms[r];
			c->local = group;
			c->host = host_group;
			rc = comm_alloc(c);
			out[r] = c;
		}
		if(rc != MH_OK) {
			for(int q = 0; q <= r; ++q) { mh_comm_destroy(out[q]); out[q] = nullptr; }
			for(int q = r + 1; q < world; ++q) if(comms[q]) nccl_api().CommDestroy(comms[q]);
			return rc;
		}
	}
	if(group) {   // peers read each other's buffers directly where the hardware allows it
		for(int a = 0; a < world; ++a)
			for(int b = 0; b < world; ++b) {
				if(dev[a] == dev[b]) continue;
				int can = 0;
				if(cudaDeviceCanAccessPeer(&can, dev[a], dev[b]) == cudaSuccess && can) {
					cudaSetDevice(dev[a]);
					if(cudaDeviceEnablePeerAccess(dev[b], 0) != cudaSuccess) cudaGetLastError();
				}
			}
	}
	return MH_OK;
}

int mh_comm_rank(const mh_comm* c) { return c ? c->rank : MH_ERR_INVALID_ARG; }
int mh_comm_world(const mh_comm* c) { return c ? c->world : MH_ERR_INVALID_ARG; }
int mh_comm_device(const mh_comm* c) { return c ? c->device : MH_ERR_INVALID_ARG; }
int mh_comm_transport(const mh_comm* c) { return !c ? MH_ERR_INVALID_ARG : (c->nccl ? 1 : (c->local ? 2 : 0)); }

void mh_comm_destroy(mh_comm* c) {
	if(!c) return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	if(c->nccl && nccl_api().ok) nccl_api().CommDestroy(c->nccl);
	release_codebook(&c->book);
	release_dectable(&c->dec);
	mh_workspace_destroy(c->ws);
	if(c->stream) cudaStreamDestroy(c->stream);
	if(c->side) cudaStreamDestroy(c->side);
	if(c->ev_side) cudaEventDestroy(c->ev_side);
	void* dptrs[] = {c->d_msg, c->d_gather, c->d_total, c->d_bits, c->d_layout, c->d_result, c->d_seams, c->d_halo, c->d_halos, c->d_hin, c->d_hlocal, c->d_hout};
	for(void* p : dptrs)
		if(p) cudaFree(p);
	if(c->h_total) cudaFreeHost(c->h_total);
	if(c->h_small) cudaFreeHost(c->h_small);
	for(auto& e : c->ev)
		if(e) cudaEventDestroy(e);
	delete c;
}

int mh_comm_reserve(mh_comm* c, uint64_t max_shard_bytes, uint64_t max_payload_bytes) {
	if(!c) return MH_ERR_INVALID_ARG;
	if(c->ws && c->ws_input >= max_shard_bytes && c->ws_payload >= max_payload_bytes) return MH_OK;
	MH_CUDA(cudaSetDevice(c->device));
	mh_workspace_destroy(c->ws);
	c->ws = nullptr;
	c->ws_input = max_shard_bytes > c->ws_input ? max_shard_bytes : c->ws_input;
	c->ws_payload = max_payload_bytes > c->ws_payload ? max_payload_bytes : c->ws_payload;
	return mh_workspace_create(c->ws_input, c->ws_payload, &c->ws);
}

uint64_t mh_shard_local_bytes(uint64_t max_payload_bytes) { return uint64_t(kShardPad) + max_payload_bytes + kHaloHead + 64; }
uint32_t mh_shard_payload_offset(void) { return kShardPad; }

int mh_comm_stats(mh_comm* c, double* out, int n, int reset) {
	if(!c || (!out && n)) return MH_ERR_INVALID_ARG;
	for(int i = 0; i < n; ++i) out[i] = i < kStatCount ? c->stats[i] : 0.0;
	if(reset) for(double& v : c->stats) v = 0.0;
	return MH_OK;
}

int mh_sharded_compress(mh_comm* c, const uint8_t* d_in, uint64_t n, int order, uint8_t* d_local, uint64_t local_cap,
                        mh_shard_layout* layout, mh_table** table_out, int prepare_decode, mh_stream_t stream) {
	if(!c || !layout || !d_local || (!d_in && n) || (order != 0 && order != 1)) return MH_ERR_INVALID_ARG;
	if(local_cap < mh_shard_local_bytes(64)) return MH_ERR_CAPACITY;
	if(reinterpret_cast<uint64_t>(d_local) & 15) return MH_ERR_INVALID_ARG;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	MH_CUDA(cudaSetDevice(c->device));
	const uint64_t pay_cap = local_cap - kShardPad - kHaloHead - 64;
	int rc = mh_comm_reserve(c, n, pay_cap);   // a no-op once sized
	if(rc != MH_OK) return rc;
	const uint32_t world = uint32_t(c->world), bins = order ? 65536u : 256u;
	// 1. local counts + edge bytes, one all-gather
	rc = launch_histogram(d_in, n, MH_PREV0, order, c->d_msg, c->ws, st);
	if(rc != MH_OK) return rc;
	shard_edges_kernel<<<1, 1, 0, st>>>(d_in, n, c->d_msg + 65536);
	count_launch(1);
	rc = all_gather(c, c->d_msg, c->d_gather, kMsgWords * sizeof(unsigned long long), st, 0);
	if(rc != MH_OK) return rc;
	// 2. the sum, with the seam pairs where they belong; to the host for the trees
	shard_sum_kernel<<<(bins + 255) / 256, 256, 0, st>>>(c->d_gather, world, bins, c->d_total);
	shard_seam_kernel<<<1, 1, 0, st>>>(c->d_gather, world, order, c->d_total, c->d_total + 65536);
	count_launch(2);
	MH_CUDA(cudaGetLastError());
	// 3. The encoder's tables are built on the device from the summed counts (mh_tables.cu) and every shard's payload
	//    size and bit offset follow from the gathered counts x code lengths: nothing between the gather and the encoder
	//    waits for the host. The counts travel to the host on a side stream meanwhile; the host builds its (identical)
	//    table from them while the encoder runs — the table file, the decoder's tables and the layout need it, not the encoder.
	if(!c->side) MH_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
	if(!c->ev_side) MH_CUDA(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
	MH_CUDA(cudaEventRecord(c->ev_side, st));
	MH_CUDA(cudaStreamWaitEvent(c->side, c->ev_side, 0));
	MH_CUDA(cudaMemcpyAsync(c->h_total, c->d_total, bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->side));
	MH_CUDA(cudaMemcpyAsync(c->h_total + 65536, c->d_total + 65536, world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->side));
	rc = launch_build_codebook(c->d_total, order, &c->book, st);
	if(rc != MH_OK) return rc;
	shard_bits_kernel<<<world, 256, 0, st>>>(c->d_gather, reinterpret_cast<const unsigned long long*>(c->book.d_enc), bins, c->d_bits);
	shard_scan_kernel<<<1, 32, 0, st>>>(c->d_bits, world, c->d_layout);
	count_launch(2);
	rc = launch_encode(d_in, n, MH_PREV0, &c->book, 0, d_local + kShardPad, pay_cap, c->d_result, c->ws, st, c->d_layout + c->rank, c->d_total + 65536 + c->rank);
	if(rc != MH_OK) return rc;
	unsigned long long* h_layout = c->h_small;
	unsigned long long* h_result = c->h_small + 2 * world;
	MH_CUDA(cudaMemcpyAsync(h_layout, c->d_layout, 2 * world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	MH_CUDA(cudaMemcpyAsync(h_result, c->d_result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	MH_CUDA(cudaStreamSynchronize(c->side));   // the counts have arrived (the encoder is still running)
	add_elapsed(c, kStatGather, 0);
	const double t0 = now_us();
	mh_table* t = nullptr;
	rc = mh_table_from_counts(reinterpret_cast<const uint64_t*>(c->h_total), order, &t);   // identical on every rank
	if(rc != MH_OK) { cudaStreamSynchronize(st); return rc; }
	c->stats[kStatTrees] += now_us() - t0;
	if(prepare_decode) {   // the decoder's tables are flattened on the host while the encoder runs
		const double t3 = now_us();
		rc = upload_dectable_for(t, &c->dec, st);
		c->stats[kStatDectable] += now_us() - t3;
		if(rc != MH_OK) { mh_table_destroy(t); return rc; }
		c->dec_serial = table_serial(t);
	}
	MH_CUDA(cudaStreamSynchronize(st));
	if(h_result[3]) {
		// the device-built tables did not fit the encoder's launch (many live contexts, very long codewords): the host's
		// tables, the classic way. Every rank sees the same tables and takes this branch together.
		const double t1 = now_us();
		rc = upload_codebook_for(t, &c->book, st);
		c->stats[kStatCodebook] += now_us() - t1;
		if(rc == MH_OK) {
			shard_bits_kernel<<<world, 256, 0, st>>>(c->d_gather, reinterpret_cast<const unsigned long long*>(c->book.d_enc), bins, c->d_bits);
			shard_scan_kernel<<<1, 32, 0, st>>>(c->d_bits, world, c->d_layout);
			count_launch(2);
			rc = launch_encode(d_in, n, uint8_t(c->h_total[65536 + c->rank]), &c->book, 0, d_local + kShardPad, pay_cap, c->d_result, c->ws, st, c->d_layout + c->rank);
		}
		if(rc != MH_OK) { mh_table_destroy(t); return rc; }
		MH_CUDA(cudaMemcpyAsync(h_layout, c->d_layout, 2 * world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		MH_CUDA(cudaMemcpyAsync(h_result, c->d_result, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		MH_CUDA(cudaStreamSynchronize(st));
	}
	MH_CUDA(cudaStreamSynchronize(st));
	c->stats[kStatCalls] += 1;
	if(h_result[2]) { mh_table_destroy(t); return MH_ERR_CAPACITY; }
	memset(layout, 0, sizeof *layout);
	layout->world = int(world);
	layout->order = order;
	layout->exact = 1;
	for(uint32_t r = 0; r < world; ++r) {
		layout->bit_base[r] = h_layout[r];
		layout->n_bits[r] = h_layout[world + r];
		layout->prev0[r] = uint8_t(c->h_total[65536 + r]);
		layout->total_bits += h_layout[world + r];
	}
	layout->dropped = h_result[1];
	if(h_result[0] != layout->n_bits[c->rank]) {   // the encoder and sum(count x length) must agree to the bit
		mh_table_destroy(t);
		set_last_error("sharded compress: the encoder's bit count differs from sum(count x code length)");
		return MH_ERR_CORRUPT_STREAM;
	}
	if(table_out) *table_out = t;
	else mh_table_destroy(t);
	return MH_OK;
}

int mh_sharded_decompress(mh_comm* c, const mh_table* t, uint8_t* d_local, uint64_t local_cap, const mh_shard_layout* layout, int speculative,
                          uint8_t* d_out, uint64_t out_capacity, uint64_t* n_out, uint64_t* out_offset, mh_stream_t stream) {
	if(!c || !t || !d_local || !layout || !n_out || layout->world != c->world) return MH_ERR_INVALID_ARG;
	if(reinterpret_cast<uint64_t>(d_local) & 15) return MH_ERR_INVALID_ARG;
	if(layout->order != table_order(t)) return MH_ERR_TYPE_MISMATCH;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	MH_CUDA(cudaSetDevice(c->device));
	const uint32_t world = uint32_t(c->world), r = uint32_t(c->rank);
	const uint64_t bits = layout->n_bits[r], base = layout->bit_base[r];
	const uint32_t phase = uint32_t(base & 7);
	const uint64_t nbytes = (phase + bits + 7) / 8;
	if(local_cap < mh_shard_local_bytes(nbytes)) return MH_ERR_CAPACITY;
	int rc = mh_comm_reserve(c, 0, nbytes + kHaloTail + kHaloHead + 64);
	if(rc != MH_OK) return rc;
	if(c->dec_serial != table_serial(t)) {
		const double t3 = now_us();
		rc = upload_dectable_for(t, &c->dec, st);
		c->stats[kStatDectable] += now_us() - t3;
		if(rc != MH_OK) return rc;
		c->dec_serial = table_serial(t);
	}
	uint8_t* pay = d_local + kShardPad;
	const bool spec = speculative && world > 1;
	if(spec) {
		// the warm-up reaches kWarmBits into the predecessor: every shard must be at least that long
		for(uint32_t g = 0; g < world; ++g)
			if(layout->n_bits[g] < uint64_t(kHaloTail) * 8) { set_last_error("sharded decompress: a shard is shorter than the warm-up; use exact mode or fewer shards"); return MH_ERR_INVALID_ARG; }
		halo_pack_kernel<<<1, 256, 0, st>>>(pay, nbytes, c->d_halo);
		count_launch(1);
		rc = all_gather(c, c->d_halo, c->d_halos, kHaloTail + kHaloHead, st, 2);
		if(rc != MH_OK) return rc;
		const uint32_t next_phase = r + 1 < world ? uint32_t(layout->bit_base[r + 1] & 7) : 0u;
		halo_splice_kernel<<<1, 256, 0, st>>>(d_local, c->d_halos, r, world, phase, next_phase, nbytes);
		count_launch(1);
		MH_CUDA(cudaGetLastError());
	}
	std::vector<uint64_t> counts(world);
	rc = decode_with_handshake(c, d_local, uint64_t(kShardPad) * 8 + phase, bits, uint64_t(kShardPad) + nbytes + (spec ? kHaloHead : 0), spec,
	                           layout->prev0[r], d_out, out_capacity, st, spec, counts.data());
	if(rc != MH_OK) return rc;
	c->stats[kStatCalls] += 1;
	uint64_t before = 0;
	for(uint32_t g = 0; g < r; ++g) before += counts[g];
	*n_out = counts[r];
	if(out_offset) *out_offset = before;
	return MH_OK;
}

// ---- ranks of ONE process on shared host buffers (the multi-GPU path of a command-line driver) ---------------------
// Every rank (its own host thread) passes the SAME host pointers; rank r takes its share. The calls return on every
// rank when the whole result is in `out`.
int mh_sharded_compress_host(mh_comm* c, const uint8_t* in, uint64_t n, int order, uint8_t* out, uint64_t out_capacity, uint64_t* out_len,
                             mh_table** table_out) {
	if(!c || !c->host || !out || !out_len || (!in && n) || (order != 0 && order != 1) || out_capacity < 1) return MH_ERR_INVALID_ARG;
	MH_CUDA(cudaSetDevice(c->device));
	LocalGroup& g = *c->host;
	const uint64_t W = uint64_t(c->world), r = uint64_t(c->rank);
	const uint64_t lo = n / W * r + (r < n % W ? r : n % W), len = n / W + (r < n % W ? 1 : 0);   // contiguous byte ranges in rank order
	if(!c->stream) MH_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	const uint64_t local_cap = mh_shard_local_bytes(len + len / 8 + 4096);
	int rc = ensure_device(&c->d_hin, &c->hin_cap, len + 64);
	if(rc == MH_OK) rc = ensure_device(&c->d_hlocal, &c->hlocal_cap, local_cap);
	mh_shard_layout layout;
	mh_table* t = nullptr;
	if(rc == MH_OK && len) rc = cuda_fail_if(cudaMemcpyAsync(c->d_hin, in + lo, len, cudaMemcpyHostToDevice, c->stream), "H2D of the shard");
	g.status[r] = rc;
	g.barrier();   // nobody enters the collective unless everybody can
	for(uint64_t q = 0; q < W; ++q)
		if(g.status[q] != MH_OK) return g.status[q];
	rc = mh_sharded_compress(c, c->d_hin, len, order, c->d_hlocal, c->hlocal_cap, &layout, &t, 0, c->stream);
	if(rc != MH_OK) return rc;
	const uint64_t total = 1 + (layout.total_bits + 7) / 8;
	*out_len = total;
	if(total > out_capacity) { mh_table_destroy(t); return MH_ERR_CAPACITY; }   // the same decision on every rank
	const uint64_t base = layout.bit_base[r], bits = layout.n_bits[r];
	const uint32_t phase = uint32_t(base & 7);
	const uint64_t nbytes = (phase + bits + 7) / 8, first = 1 + (base >> 3);
	const uint8_t* pay = c->d_hlocal + kShardPad;
	cudaError_t ce = cudaSuccess;
	g.seam[r] = 0;
	g.value[r] = 0;
	if(nbytes) {
		if(phase) {   // my first byte is my predecessor's last: it keeps that one's bits; mine are merged below
			ce = cudaMemcpyAsync(&g.seam[r], pay, 1, cudaMemcpyDeviceToHost, c->stream);
			if(ce == cudaSuccess && nbytes > 1) ce = cudaMemcpyAsync(out + first + 1, pay + 1, nbytes - 1, cudaMemcpyDeviceToHost, c->stream);
			g.value[r] = (first << 8) | phase;
		} else {
			ce = cudaMemcpyAsync(out + first, pay, nbytes, cudaMemcpyDeviceToHost, c->stream);
		}
	}
	if(ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
	g.status[r] = ce == cudaSuccess ? MH_OK : cuda_fail(ce, "D2H of the payload shard");
	g.barrier();
	rc = MH_OK;
	for(uint64_t q = 0; q < W; ++q)
		if(g.status[q] != MH_OK) rc = g.status[q];
	if(r == 0 && rc == MH_OK) {
		for(uint64_t q = 0; q < W; ++q)
			if(g.value[q]) out[g.value[q] >> 8] = uint8_t(out[g.value[q] >> 8] | (g.seam[q] & (0xFFu >> (g.value[q] & 7))));
		out[0] = uint8_t(0x30 | ((~order & 1) << 3) | ((8 - layout.total_bits % 8) % 8));   // src/coding.cpp:88
	}
	g.barrier();   // the stream is complete when anybody returns
	if(rc == MH_OK && table_out && r == 0) *table_out = t;
	else mh_table_destroy(t);
	return rc;
}

int mh_sharded_decompress_host(mh_comm* c, const mh_table* t, const uint8_t* stream, uint64_t stream_len, uint8_t* out, uint64_t out_capacity,
                               uint64_t* out_len) {
	if(!c || !c->host || !t || !stream || !out_len) return MH_ERR_INVALID_ARG;
	if(stream_len < 1) return MH_ERR_BAD_HEADER;
	const uint8_t header = stream[0];
	if((header & 0xF0) != 0x30) return MH_ERR_BAD_HEADER;                       // src/coding.cpp:103-106
	if(((~header >> 3) & 1) != table_order(t)) return MH_ERR_TYPE_MISMATCH;     // src/coding.cpp:107-110
	const uint8_t* payload = stream + 1;
	const uint64_t payload_bytes = stream_len - 1, remainder = header & 7;
	const uint64_t B = payload_bytes * 8 < remainder ? 0 : payload_bytes * 8 - remainder;
	const uint64_t W = uint64_t(c->world), r = uint64_t(c->rank);
	if(W > 1 && B / W < 4ull * kHaloTail * 8) { set_last_error("sharded decompress: the stream is too short to cut; use one GPU"); return MH_ERR_INVALID_ARG; }
	MH_CUDA(cudaSetDevice(c->device));
	LocalGroup& g = *c->host;
	if(!c->stream) MH_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	// bit ranges: cuts at multiples of 32 bits, wherever they fall inside codewords
	auto cut = [&](uint64_t q) { return q >= W ? B : (B / W * q) & ~uint64_t(31); };
	const uint64_t b0 = cut(r), b1 = cut(r + 1);
	const uint64_t s0 = r == 0 ? 0 : (((b0 - kWarmBits) >> 3) - 8) & ~uint64_t(15);          // my slice of the payload: warm-up in front ...
	const uint64_t e0 = (b1 >> 3) + kHaloHead + 8 < payload_bytes ? (b1 >> 3) + kHaloHead + 8 : payload_bytes;   // ... room for the last codeword behind
	int rc = ensure_device(&c->d_hlocal, &c->hlocal_cap, e0 - s0 + 64);
	if(rc == MH_OK) rc = ensure_device(&c->d_hout, &c->hout_cap, (e0 - s0) * 3 + 4096);
	if(rc == MH_OK) rc = mh_comm_reserve(c, 0, e0 - s0 + 64);
	if(rc == MH_OK && table_serial(t) != c->dec_serial) {
		rc = upload_dectable_for(t, &c->dec, c->stream);
		if(rc == MH_OK) c->dec_serial = table_serial(t);
	}
	if(rc == MH_OK) rc = cuda_fail_if(cudaMemsetAsync(c->d_hlocal + (e0 - s0), 0, 64, c->stream), "clearing the slack behind the slice");
	if(rc == MH_OK && e0 > s0) rc = cuda_fail_if(cudaMemcpyAsync(c->d_hlocal, payload + s0, e0 - s0, cudaMemcpyHostToDevice, c->stream), "H2D of the slice");
	std::vector<uint64_t> counts(W);
	for(int attempt = 0;; ++attempt) {
		g.status[r] = rc;
		g.barrier();   // nobody enters the collective unless everybody can
		for(uint64_t q = 0; q < W; ++q)
			if(g.status[q] != MH_OK) return g.status[q];
		rc = decode_with_handshake(c, c->d_hlocal, b0 - 8 * s0, b1 - b0, e0 - s0, W > 1, MH_PREV0, c->d_hout, c->hout_cap, c->stream, false, counts.data());
		if(rc != MH_ERR_CAPACITY || attempt >= 2) break;
		// some rank's symbols outgrew its buffer (the count pass told every rank how many each has): grow mine if it was me
		rc = counts[r] > c->hout_cap ? ensure_device(&c->d_hout, &c->hout_cap, counts[r] + 64) : MH_OK;
	}
	if(rc != MH_OK) return rc;
	uint64_t before = 0, total = 0;
	for(uint64_t q = 0; q < W; ++q) {
		if(q < r) before += counts[q];
		total += counts[q];
	}
	*out_len = total;
	if(total > out_capacity || !out) return total > out_capacity || total ? MH_ERR_CAPACITY : MH_OK;
	cudaError_t ce = counts[r] ? cudaMemcpyAsync(out + before, c->d_hout, counts[r], cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
	if(ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
	g.status[r] = ce == cudaSuccess ? MH_OK : cuda_fail(ce, "D2H of the decoded bytes");
	g.barrier();
	for(uint64_t q = 0; q < W; ++q)
		if(g.status[q] != MH_OK) return g.status[q];
	return MH_OK;
}

}  // extern "C"
