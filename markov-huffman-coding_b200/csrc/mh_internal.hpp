// mh_internal.hpp — shared declarations between the C-ABI glue (mh_api.cu) and the kernel files.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>

#include "../../include/mh_gpu.h"

namespace mh {

extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(uint64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

// Brackets one kernel launch with CUDA events on its stream while profiling is enabled (mh_profile_enable).
struct ProfScope {
	ProfScope(const char* name, cudaStream_t st);
	~ProfScope();
	int slot;
	cudaStream_t st;
	cudaEvent_t end = nullptr;
};

int cuda_fail(cudaError_t e, const char* what);   // records mh_last_error(), returns MH_ERR_CUDA
#define MH_CUDA(call)                                              \
	do {                                                           \
		cudaError_t e_ = (call);                                   \
		if(e_ != cudaSuccess) return ::mh::cuda_fail(e_, #call);   \
	} while(0)

int sm_count();          // multiprocessors of the CURRENT device (cached per device)
int max_smem_optin();    // bytes of opt-in dynamic shared memory per block on the current device (cached per device)
// Function attributes (opt-in shared memory) are per device: true exactly once per (mask, current device).
bool first_use_on_device(std::atomic<uint64_t>& done_mask);

// Tunables: experiment / test overrides. Read from the environment ONCE when the library is loaded (MH_ENC_FMT,
// MH_DEC_SUB_BITS_MARKOV, MH_DEC_SUB_BITS_HUFFMAN, MH_DEC_PAIR, MH_DEC_WRITE_THREADS, MH_PIPE_MIN_BYTES,
// MH_PIPE_CHUNK_BYTES, MH_ENC_TMA, MH_ENC_WARP); afterwards only mh_tunable_set changes them. -1 = the documented default.
enum Tunable { kTunEncFmt = 0, kTunDecSubBitsMarkov, kTunDecSubBitsHuffman, kTunDecPair, kTunDecWriteThreads, kTunPipeMinBytes,
               kTunPipeChunkBytes, kTunEncPipeChunkBytes, kTunEncTma, kTunDecCpGeo, kTunEncWarp, kTunEncSpt, kTunCount };
long long tunable(Tunable t);

// ---- tunables (env overrides exist for experiments; defaults are what DESIGN.md documents) ------------
constexpr int kEncThreads = 480;                      // worker threads per CTA: a tile is kEncThreads x SPT input bytes
constexpr int kEncCtaThreads = kEncThreads + 32;      // + the scanner warp (decoupled look-back, one iteration ahead of the workers)
constexpr int kEncStageMaxWords = 14336;              // at most 56 KiB of staged output bits per tile
constexpr int kEncBoxSmemLimit = 72 * 1024;           // largest u32 box table staged in shared memory (R <= 135)
constexpr int kEncBoxMaxBits = 27;                    // u32 entry: 5-bit length | 27-bit right-aligned code
constexpr int kEncCtxMaxBits = 29;                    // context-row entry: 5-bit length | 8-bit next row | 16-bit code; longer codewords (<= 29: 480 x 32 symbols of them still fit the staging area) escape to the wide table
constexpr int kEncCtxMaxRows = 96;                    // live contexts + null row; table + staging must leave room for 2 CTAs/SM
constexpr int kEncCtxSmemLimit = 112 * 1024;          // table + staging area of one CTA
constexpr int kDecThreads = 1024;                     // subsequences per chunk (one thread each)
constexpr int kDecWriteMaxThreads = 1024;             // D4 threads per CTA, as many as table + 192 B of rings per thread leave room for (text: 768)
constexpr int kDecMinSubBits = 256;
constexpr uint32_t kDecMaxSubBitsMarkov = 8192;       // measured best of 2048..16384 on the 1 GiB Markov text
constexpr uint32_t kDecMaxSubBitsHuffman = 8192;      // 1 GiB text with -h: D1 1.48 ms at 2048, 1.23 at 4096, 1.05 at 8192
constexpr uint64_t kDecTargetSubs = 80000;            // fewest subsequences worth having: measured best of 1024..8192 bits at 30 MB, 100 MB, 300 MB, 1 GiB
constexpr int kDecPairBytes = 64 * 1024 + 256 + 64 + 64 * 256;   // pair table: <= 64 rows x 256 x u32, then rank[256], live[64], len1[<= 64 x 256]
constexpr int kDecWarmSubs = 8;                       // overlap subsequences re-decoded by the next chunk

// ---- encode look-back descriptors -----------------------------------------------------------------------
// status word: [63:62] 0 = empty, 1 = aggregate (this tile only), 2 = inclusive (all tiles up to this one);
//              [61:0]  bit count. The matching tail word holds the last min(bits, 31) bits, right-aligned.
constexpr uint64_t kDescAggregate = 1ull << 62;
constexpr uint64_t kDescInclusive = 2ull << 62;
constexpr uint64_t kDescValueMask = (1ull << 62) - 1;

}  // namespace mh

// Device scratch. All pointers are device memory owned by the workspace.
struct mh_workspace {
	// encode
	uint64_t* enc_desc = nullptr;         // enc_tiles_cap x (u64 aggregate, u64 inclusive, u32 tail) as three arrays
	uint64_t enc_tiles_cap = 0;
	uint32_t* counters = nullptr;         // [16] dynamic tile counters / flags
	// histogram
	uint32_t* hist_params = nullptr;      // [8] lo, range, replicas, ...
	// decode
	uint32_t* dec_state = nullptr;        // [dec_subs_cap] end state of each subsequence: rel_bits << 8 | context
	uint32_t* dec_count = nullptr;        // [dec_subs_cap] symbols that start in each subsequence
	uint32_t* dec_prefix = nullptr;       // [dec_subs_cap] symbols of the same chunk before each subsequence
	uint64_t dec_subs_cap = 0;
	uint32_t* dec_seam = nullptr;         // [dec_chunks_cap] boundary state as seen by the next chunk's warm-up
	uint64_t* dec_chunk_total = nullptr;  // [dec_chunks_cap]
	uint64_t* dec_chunk_base = nullptr;   // [dec_chunks_cap + 1]
	uint64_t dec_chunks_cap = 0;
	uint32_t* dec_flags = nullptr;        // [8] 0: seam mismatch pending, 1: error status, ...
};

struct mh_codebook {
	uint64_t* d_enc = nullptr;     // [ntab * 256] len << 56 | code
	uint32_t* d_box = nullptr;     // [(R + 1)^2] with a zero border (order 1) or [256] (order 0): len << 27 | code
	uint32_t box_lo = 0, box_r = 256;
	bool has_box = false;
	uint32_t* d_ctx = nullptr;     // [ctx_rows * 256] len << 27 | next row << 16 | code (max_bits <= 16, few live contexts)
	uint32_t* h_ctx = nullptr;     // pinned
	uint32_t ctx_rows = 0;         // 0: no context-row table
	uint64_t* h_stage = nullptr;   // pinned image the async upload reads from
	uint32_t* h_box = nullptr;     // pinned
	cudaEvent_t uploaded = nullptr;
	int order = 1;
	int max_bits = 0;
	// tables built on the device from the device-resident histogram (mh_tables.cu): the host knows neither the number of
	// context rows nor the longest codeword; the encoder reads them from d_meta and reports when its launch did not fit
	uint32_t* d_meta = nullptr;    // [8] rows, status, longest codeword, live contexts; then rank[256] bytes
	bool device_built = false;
};

struct mh_dectable {
	uint16_t* d_lut = nullptr;     // [ntab * 256]
	uint32_t* d_walk = nullptr;    // [ntab * 512]
	uint16_t* h_lut = nullptr;     // pinned
	uint32_t* h_walk = nullptr;    // pinned
	uint32_t* d_pair = nullptr;    // [pair_rows * 256] two-symbol entries, then rank[256] + live[64] bytes (see flatten_pairlut)
	uint32_t* h_pair = nullptr;    // pinned
	uint32_t pair_rows = 0;        // context rows + null row + prefix rows; 0: no pair table (more than 63 live contexts)
	uint32_t pair_ctx_rows = 0;    // context rows (= index of the null row)
	cudaEvent_t uploaded = nullptr;
	int order = 1;
	int max_bits = 0;
};

namespace mh {

// accumulate: add to d_counts instead of overwriting it (chunked inputs: the counts of the chunks add up)
int launch_histogram(const uint8_t* d_in, uint64_t n, uint8_t prev0, int order, unsigned long long* d_counts,
                     mh_workspace* ws, cudaStream_t st, bool accumulate = false);
// d_prev0 (optional): the byte before d_in[0] in DEVICE memory (replaces prev0).
// d_bit_base (optional): the shard's global bit offset in DEVICE memory — its low three bits are the bit phase of the
// first codeword and replace `bit_base` (a sharded compress computes the offsets on the device and never waits for them)
int launch_encode(const uint8_t* d_in, uint64_t n, uint8_t prev0, const mh_codebook* cb, uint64_t bit_base,
                  uint8_t* d_out, uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws, cudaStream_t st,
                  const unsigned long long* d_bit_base = nullptr, const unsigned long long* d_prev0 = nullptr);
int launch_decode(const uint8_t* d_bits, uint64_t bit_base, uint64_t n_bits, uint8_t prev0, const mh_dectable* dt, uint8_t* d_out,
                  uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws, cudaStream_t st, int fix_iters);
int launch_decode_shard(const uint8_t* d_bits, uint32_t start_bit, uint64_t n_bits, uint64_t buf_bytes, int exact_start, uint8_t prev0,
                        uint32_t warm_bits, int stream_end, const mh_dectable* dt, uint8_t* d_out, uint64_t out_capacity,
                        unsigned long long* d_result, mh_workspace* ws, cudaStream_t st, int fix_iters);
int launch_build_codebook(const unsigned long long* d_counts, int order, mh_codebook* cb, cudaStream_t st);   // mh_tables.cu
uint32_t decode_sub_bits(int order, uint64_t n_bits);   // subsequence size in bits for a stream of n_bits of this coder type
uint64_t decode_max_subs(uint64_t max_payload_bytes);   // workspace bound on the number of subsequences
uint64_t encode_tiles_for(uint64_t n);               // worst-case tile count for n input bytes

}  // namespace mh
