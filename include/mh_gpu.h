/* include/mh_gpu.h — C ABI of the B200-native Markov-Huffman codec (libmh_gpu.so).
 *
 * This is the drop-in boundary for the reference's hot path (jeremy-rifkin/Markov-Huffman-Coding). The reference
 * has no FFI; its seam is the abstract class i_coding_provider (src/coding.h:18-35) plus the free function
 * construct_table (src/main.cpp:29). Each entry point below names the reference interface it replaces.
 * Plain pointers and sizes only: no CUDA, torch or C++ types cross this boundary (a CUDA stream is passed as
 * void*). Every function returns MH_OK (0) or a negative mh_status; nothing exits or throws.
 *
 * There is no CPU fallback: every mh_gpu_* / mh_session_* call runs hand-written sm_100a kernels and fails with
 * MH_ERR_CUDA / MH_ERR_NO_DEVICE when no device is usable.
 */
#ifndef MH_GPU_H
#define MH_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mh_status {
	MH_OK = 0,
	MH_ERR_INVALID_ARG = -1,
	MH_ERR_CUDA = -2,              /* a CUDA runtime call failed; see mh_last_error() */
	MH_ERR_NO_DEVICE = -3,
	MH_ERR_CAPACITY = -4,          /* an output buffer is too small */
	MH_ERR_BAD_TABLE = -5,         /* table file truncated / malformed */
	MH_ERR_CODE_TOO_LONG = -6,     /* a codeword exceeds MH_MAX_CODE_BITS (device tables only) */
	MH_ERR_BAD_HEADER = -7,        /* "Input appears corrupt"                      (src/coding.cpp:103-106) */
	MH_ERR_TYPE_MISMATCH = -8,     /* "File encoding method does not match ..."    (src/coding.cpp:107-110) */
	MH_ERR_CORRUPT_STREAM = -9,    /* decode reached a null table entry / ran past the payload */
	MH_ERR_COUNT_WRAPPED = -10,    /* a live count is a multiple of 2^32: the reference's int counter is 0 (F3) */
	MH_ERR_NOT_CONVERGED = -11,    /* decode seam fix-up needs more iterations (device API only) */
	MH_ERR_WORKSPACE = -12         /* workspace too small for this call */
} mh_status;

#define MH_MAX_CODE_BITS 56        /* device codebook entry: 8-bit length + 56-bit right-aligned code */
#define MH_ORDER_HUFFMAN 0         /* -h : one tree, get_type() == 0 (src/huffman.cpp:48-50) */
#define MH_ORDER_MARKOV 1          /* default: 256 trees indexed by the previous byte, get_type() == 1 */
#define MH_PREV0 0x20              /* the reference seeds prev with ' ' (src/main.cpp:32, src/coding.cpp:67,118) */

typedef void* mh_stream_t;         /* a cudaStream_t, or NULL for the default stream */

const char* mh_status_string(int status);
const char* mh_last_error(void);   /* thread-local detail of the last MH_ERR_CUDA */
int mh_device_count(void);
/* Free and total memory of a device in bytes (cudaMemGetInfo): a host driver sizes its session from it and lets the
 * session stream larger files through in chunks. */
int mh_device_memory(int device, uint64_t* free_bytes, uint64_t* total_bytes);
int mh_version(void);
/* Experiment / test overrides of the library's tunables ("enc_fmt", "dec_sub_bits_markov", "dec_sub_bits_huffman",
 * "dec_pair", "dec_write_threads", "pipe_min_bytes", "pipe_chunk_bytes", "enc_pipe_chunk_bytes", "enc_tma", "dec_cp_geo",
 * "enc_warp" — 1 selects the encoder with warp-private tiles —, "enc_spt" — 32 switches the optimistic 64-symbols-per-thread
 * encoder launch for device-built tables off —); -1 = the documented default. The matching environment variables (MH_ENC_FMT, ...) are read once, when the library is loaded. */
int mh_tunable_set(const char* name, long long value);
int mh_tunable_get(const char* name, long long* value);
/* Page-locked host memory for the host-buffer calls below: copies from / to pageable memory cannot overlap with the
 * kernels (a host driver reads its file straight into such a buffer). NULL when the allocation fails. */
void* mh_pinned_alloc(size_t bytes);
void mh_pinned_free(void* p);

/* ---------------------------------------------------------------------------------------------------------
 * Host side: the coding tables. Replaces huffman_table / markov_huffman_table (src/huffman.{h,cpp},
 * src/markov_huffman.{h,cpp}), tree_node (src/tree.h) and min_pq (src/min_pq.tpp). Microsecond work that
 * stays on the host and reproduces the reference's heap tie-breaking and int32 weight arithmetic exactly.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct mh_table mh_table;

/* huffman_table(int*) (src/huffman.cpp:18-20) / markov_huffman_table(int*) (src/markov_huffman.cpp:9-13).
 * counts: 256 (order 0) or 65536 (order 1, index 256*prev + c) 64-bit counts as mh_gpu_histogram produces them;
 * they are truncated to the reference's int32 before use. */
int mh_table_from_counts(const uint64_t* counts, int order, mh_table** out);
/* huffman_table(bitbuffer&) / markov_huffman_table(bitbuffer&) (src/huffman.cpp:22-25, src/markov_huffman.cpp:15-25):
 * load an encoding-table file image. The first bit selects the kind (src/main.cpp:147-161). */
int mh_table_from_bytes(const uint8_t* bytes, size_t n, mh_table** out);
/* write_coding_tree (src/coding.h:23; src/huffman.cpp:83-85,174-188; src/markov_huffman.cpp:80-88). */
int mh_table_serialize(const mh_table* t, uint8_t* out, size_t cap, size_t* n_out);
/* get_type() (src/coding.h:32). */
int mh_table_order(const mh_table* t);
/* huffman_table::empty() for context prev (src/huffman.cpp:44-46); prev is ignored for order 0. */
int mh_table_context_empty(const mh_table* t, int prev);
/* get_encoding(prev, c) (src/coding.h:33): *len = bit length (0: no codeword), bits = MSB-first bytes. */
int mh_table_code(const mh_table* t, int prev, int c, uint8_t bits[32], int* len);
int mh_table_max_code_bits(const mh_table* t);
/* All code lengths at once: lens[256*prev + c] for order 1 (65536 entries), lens[c] for order 0 (256 entries).
 * With the local histogram this gives a shard's payload size without touching the data (sum of count x length). */
int mh_table_code_lengths(const mh_table* t, uint8_t* lens, size_t cap);
/* decoding_lookup(prev, w) (src/coding.h:34): kind 0 null, 1 leaf (value, depth), 2 internal node at depth 8. */
int mh_table_lookup(const mh_table* t, int prev, int window, int* kind, int* value, int* depth);
/* The decoder's two-symbol table over the live contexts (derived from decoding_lookup, src/coding.h:34; layout in
 * csrc/mh_host.hpp, flatten_pairlut): table[rows * 256] u32 entries, maps = rank[256] | live[64] | len1[ctx_rows * 256].
 * `table` needs 64 * 256 entries, `maps` 256 + 64 * 257 bytes. *rows = 0 when the table has more than 63 live contexts
 * (the decoder then uses the 8-bit LUT alone). Introspection for tests; the GPU path builds the same image itself. */
int mh_table_pair_lut(const mh_table* t, uint32_t* table, uint8_t* maps, uint32_t* rows, uint32_t* ctx_rows);
/* print_table() + print_tree() (src/coding.h:21-22): the `-g` dump, written to the caller's buffer. */
int mh_table_debug_dump(const mh_table* t, char* out, size_t cap, size_t* n_out);
void mh_table_destroy(mh_table* t);

/* ---------------------------------------------------------------------------------------------------------
 * Device side: the three hot loops on device-resident buffers (caller owns all buffers and the stream).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct mh_codebook mh_codebook;     /* flat (prev, c) -> (code, length) table in device memory */
typedef struct mh_dectable mh_dectable;     /* 8-bit LUTs + flattened trees for the > 8-bit walk, in device memory */
typedef struct mh_workspace mh_workspace;   /* device scratch: scan descriptors, subsequence states */

int mh_codebook_create(const mh_table* t, mh_codebook** out);
/* Re-upload another table into an existing handle, stream-ordered, without allocating or blocking the host. */
int mh_codebook_update(mh_codebook* cb, const mh_table* t, mh_stream_t stream);
void mh_codebook_destroy(mh_codebook* cb);
/* The encoder's tables built ON THE DEVICE from the device-resident histogram (d_counts as mh_gpu_histogram writes it):
 * the same heap, tie-breaking and int32 weight arithmetic as the host's (src/huffman.cpp:131-164, src/min_pq.tpp), one
 * warp per context, stream-ordered, no copy to the host and no host wait between the histogram and the encoder. The
 * host's table object (for the table file and the decoder tables) can be built from a copy of the counts meanwhile. mh_gpu_encode checks
 * that its launch fits the tables that were built; when it does not (more than ~59 live contexts, a codeword longer
 * than 28 bits, a count that wrapped to 0) it writes nothing and sets d_result[3] != 0: encode again with a codebook
 * made from the host table (mh_codebook_update). mh_codebook_create_empty makes a handle without a table. */
int mh_codebook_create_empty(mh_codebook** out);
int mh_codebook_build_device(mh_codebook* cb, const uint64_t* d_counts, int order, mh_stream_t stream);
/* Test introspection: copies the handle's device tables to the host (synchronises the device). enc: 65536 x u64 (256 for
 * order 0), ctx: up to 96 x 256 x u32, meta: 8 x u32 (rows incl. the null row, status, longest codeword, live contexts). */
int mh_codebook_download(mh_codebook* cb, uint64_t* enc, uint32_t* ctx, uint32_t* meta);
int mh_dectable_create(const mh_table* t, mh_dectable** out);
int mh_dectable_update(mh_dectable* dt, const mh_table* t, mh_stream_t stream);
void mh_dectable_destroy(mh_dectable* dt);
/* Scratch for inputs up to max_input_bytes and payloads up to max_payload_bytes (either may be 0). */
int mh_workspace_create(uint64_t max_input_bytes, uint64_t max_payload_bytes, mh_workspace** out);
void mh_workspace_destroy(mh_workspace* ws);

/* Replaces construct_table + its two lambdas (src/main.cpp:29-39, :168-170, :176-178).
 * d_counts[256] (order 0) or d_counts[65536] (order 1) is OVERWRITTEN with the counts of d_in[0..n) where the
 * byte before d_in[0] is prev0 (MH_PREV0 for a whole file; the previous shard's last byte when sharding). */
int mh_gpu_histogram(const uint8_t* d_in, uint64_t n, uint8_t prev0, int order, uint64_t* d_counts,
                     mh_workspace* ws, mh_stream_t stream);

/* Replaces the loop of i_coding_provider::compress (src/coding.cpp:61-94) and the bit packer
 * (src/bitbuffer.cpp:21-73,170-180). Writes the payload (no header byte) to d_out, MSB-first. The first payload
 * bit lands at bit (bit_base & 7) of d_out[0] so that byte-range shards concatenate: shard g passes the global
 * bit offset of its first codeword and the caller ORs the seam byte. d_out needs 4-byte alignment and
 * out_capacity >= ceil(((bit_base & 7) + total_bits) / 32) * 4 bytes; bytes past the payload end inside the last
 * 32-bit word are written as zero. d_result is 4 x uint64: [0] = payload bits produced, [1] = symbols that had no
 * codeword and were dropped like the reference does (assert compiled out, src/coding.cpp:72), [2] = 1 if
 * out_capacity was too small (nothing useful was written), [3] != 0: the codebook's device-built tables do not fit this
 * encoder launch (see mh_codebook_build_device); nothing was written. */
int mh_gpu_encode(const uint8_t* d_in, uint64_t n, uint8_t prev0, const mh_codebook* cb, uint64_t bit_base,
                  uint8_t* d_out, uint64_t out_capacity, uint64_t* d_result, mh_workspace* ws, mh_stream_t stream);

/* Replaces the loop of i_coding_provider::decompress (src/coding.cpp:118-157) and the bit reader
 * (src/bitbuffer.cpp:75-140). d_bits: payload (no header byte), 4-byte aligned, n_bits payload bits starting at
 * bit (bit_base & 7) of d_bits[0] (0 for a whole stream; a byte-range shard passes the global bit offset of its
 * first codeword, as in mh_gpu_encode); prev0 is the byte decoded just before. Bits past the payload read as
 * zero (pop_rest pads, src/bitbuffer.cpp:129-140).
 * d_result is 4 x uint64: [0] = bytes decoded, [1] = 0 or a negative mh_status that stopped the write pass
 * (MH_ERR_CAPACITY, MH_ERR_NOT_CONVERGED), [2] = 0 or MH_ERR_CORRUPT_STREAM (bytes were still written, as the
 * reference would), [3] reserved. */
int mh_gpu_decode(const uint8_t* d_bits, uint64_t bit_base, uint64_t n_bits, uint8_t prev0, const mh_dectable* dt,
                  uint8_t* d_out, uint64_t out_capacity, uint64_t* d_result, mh_workspace* ws, mh_stream_t stream);

/* Decoding ONE stream on several GPUs by bit ranges. A shard decodes the bits
 * [start_bit, start_bit + n_bits) of d_bits (start_bit in 0..31; d_bits 4-byte aligned; buf_bytes readable bytes, which
 * must include a few bytes past the range so the last codeword can complete). Shard 0 knows its state
 * (exact_start = 1, prev0, warm_bits = 0); every other shard starts warm_bits (a multiple of MH_DECODE_WARM_UNIT,
 * included in n_bits) BEFORE the first bit it owns, guesses, and lets the self-synchronising decode converge over
 * that warm-up, whose symbols it does not emit. d_result is 4 x uint64: [0] symbols this shard owns (written to d_out from offset 0), [1]/[2] as
 * mh_gpu_decode, [3] = seam words: bits 63..32 the decoder state at the ownership start as the warm-up found it,
 * bits 31..0 the state at the range end; a state is (bits past the boundary << 8 | previous symbol).
 * The handshake: shard g is consistent when its start word equals shard g-1's end word; otherwise it is decoded
 * again with exact_start = 1 from that state (rare: the warm-up is several synchronisation distances long).
 * stream_end = 1 on the last shard enables the "last codeword ends exactly at the payload end" check. */
int mh_gpu_decode_shard(const uint8_t* d_bits, uint32_t start_bit, uint64_t n_bits, uint64_t buf_bytes, int exact_start,
                        uint8_t prev0, uint32_t warm_bits, int stream_end, const mh_dectable* dt, uint8_t* d_out,
                        uint64_t out_capacity, uint64_t* d_result, mh_workspace* ws, mh_stream_t stream);
#define MH_DECODE_WARM_UNIT 8192   /* every subsequence size the decoder picks divides this */
/* Subsequence size (bits) the decoder picks for a stream of n_bits of this coder type. */
uint32_t mh_decode_subsequence_bits(int order, uint64_t n_bits);

/* ---------------------------------------------------------------------------------------------------------
 * Host-buffer calls: what a maintainer binds in place of compress(FILE*, FILE*) / decompress(FILE*, FILE*)
 * (src/coding.h:26-27) and of the table construction in main (src/main.cpp:164-183). A session owns a stream,
 * pinned staging and device buffers sized at creation; calls are synchronous and include the H2D / D2H copies.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct mh_session mh_session;

int mh_session_create(int device, uint64_t max_input_bytes, mh_session** out);
/* Same with an explicit bound for the compressed side (stream bytes), e.g. for a session that only extracts: the
 * decoded size is not stored in the stream, so the uncompressed-side buffer grows on demand during decompress. */
int mh_session_create_sized(int device, uint64_t max_input_bytes, uint64_t max_stream_bytes, mh_session** out);
void mh_session_destroy(mh_session* s);

/* `markovhuffman in -o out [-h] -d table` minus the file I/O: histogram -> tables -> encode.
 * in[0..n) host bytes. out receives header byte + payload (the exact compressed file image).
 * *table_out (optional) receives the table that was built; the caller destroys it. */
int mh_session_compress(mh_session* s, const uint8_t* in, uint64_t n, int order,
                        uint8_t* out, uint64_t out_capacity, uint64_t* out_len, mh_table** table_out);
/* `markovhuffman in -o out -e table`: encode with a given table (src/main.cpp:137-162 then :211).
 * *dropped (optional) = symbols without a codeword. */
int mh_session_compress_with_table(mh_session* s, const mh_table* t, const uint8_t* in, uint64_t n,
                                   uint8_t* out, uint64_t out_capacity, uint64_t* out_len, uint64_t* dropped);
/* `markovhuffman in -o out -x -e table`: header checks (src/coding.cpp:100-116) then decode.
 * Call with out == NULL to decode on the device and get the size in *out_len; mh_session_fetch then delivers it.
 * The session's buffers are fixed at creation: a stream that is larger than the compressed-side buffer, or that
 * decodes to more bytes than the uncompressed-side buffer holds, is decoded in bit-range chunks whose bytes leave for
 * `out` chunk by chunk. Without `out` such a call only reports the size and nothing stays on the device:
 * mh_session_fetch then returns MH_ERR_WORKSPACE, and the caller repeats mh_session_decompress with `out`. */
int mh_session_decompress(mh_session* s, const mh_table* t, const uint8_t* stream, uint64_t stream_len,
                          uint8_t* out, uint64_t out_capacity, uint64_t* out_len);
/* Copies the bytes decoded by the last mh_session_decompress(out == NULL) call to the host: the size query and the
 * fetch then cost one decode, not two. MH_ERR_WORKSPACE (with *out_len = 0): that call decoded in chunks and only
 * counted — nothing is resident; call mh_session_decompress again with the output buffer. */
int mh_session_fetch(mh_session* s, uint8_t* out, uint64_t out_capacity, uint64_t* out_len);
/* Histogram only (host buffer in, host counts out): construct_table on the device. */
int mh_session_histogram(mh_session* s, const uint8_t* in, uint64_t n, int order, uint64_t* counts);

/* ---------------------------------------------------------------------------------------------------------
 * ONE logical stream over several GPUs: what a multi-GPU host driver binds in place of the callers of
 * compress / decompress (src/main.cpp:204-212). One host thread or process per GPU, each with its own mh_comm; every
 * rank makes the same mh_sharded_* calls (they are collective). The transport is NCCL over NVLink / NVSwitch
 * (libnccl.so.2, loaded at run time) or, for ranks inside one process, an in-process transport.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct mh_comm mh_comm;
#define MH_COMM_ID_BYTES 128
#define MH_MAX_SHARDS 64
int mh_comm_available(void);                        /* 1 when NCCL could be loaded */
/* One process per GPU: rank 0 makes an id, the caller hands it to every rank (its own bootstrap: MPI, a file, torch). */
int mh_comm_unique_id(uint8_t id[MH_COMM_ID_BYTES]);
int mh_comm_create(int device, int rank, int world, const uint8_t id[MH_COMM_ID_BYTES], mh_comm** out);
/* One process, `world` ranks on devices[r] (NULL: rank r on device r mod device count); out[world]. use_nccl: 0 the
 * in-process transport (several ranks may then share a device), 1 NCCL when every rank has its own device and NCCL
 * loads, else in-process, 2 NCCL or fail. Every rank's calls must then come from its own host thread. */
int mh_comm_create_local(int world, const int* devices, int use_nccl, mh_comm** out);
int mh_comm_rank(const mh_comm* c);
int mh_comm_world(const mh_comm* c);
int mh_comm_device(const mh_comm* c);
int mh_comm_transport(const mh_comm* c);            /* 0 none (world 1), 1 NCCL, 2 in-process */
void mh_comm_destroy(mh_comm* c);
/* Size the rank's scratch for shards up to max_shard_bytes / payloads up to max_payload_bytes ahead of time; the
 * mh_sharded_* calls do this themselves on first use and allocate nothing afterwards. */
int mh_comm_reserve(mh_comm* c, uint64_t max_shard_bytes, uint64_t max_payload_bytes);
/* Accumulated per-rank timings in microseconds: [0] histogram all-gather, [1] halo all-gather, [2] seam all-gather
 * (device time on the stream, waiting for the slowest rank included), [3] host tree build, [4] host codebook flatten +
 * upload, [5] host decode-table flatten + upload, [6] extra seam rounds, [7] calls. reset != 0 clears them. */
int mh_comm_stats(mh_comm* c, double* out, int n, int reset);

typedef struct mh_shard_layout {            /* how one stream is cut into `world` shards */
	int world, order;
	int exact;                              /* cuts are codeword boundaries and prev0[] is known (what compress produces) */
	uint64_t total_bits;                    /* payload bits of the whole stream */
	uint64_t dropped;                       /* this rank's symbols without a codeword (src/coding.cpp:72) */
	uint64_t bit_base[MH_MAX_SHARDS];       /* global bit offset of each shard's first bit */
	uint64_t n_bits[MH_MAX_SHARDS];
	uint8_t prev0[MH_MAX_SHARDS];           /* the byte before each shard's first symbol */
} mh_shard_layout;

/* A rank's local buffer holds its payload at d_local + mh_shard_payload_offset(), first bit at bit (bit_base & 7) of
 * that byte, with room in front and behind for the neighbours' halos; it needs mh_shard_local_bytes(payload bytes),
 * 16-byte aligned. The whole stream is the shards' payloads concatenated at their bit offsets (seam bytes OR-merged),
 * after the header byte 0 0 1 1 E R R R. */
uint64_t mh_shard_local_bytes(uint64_t max_payload_bytes);
uint32_t mh_shard_payload_offset(void);

/* Collective. Rank r holds bytes d_in[0..n) of the logical input (device memory; byte ranges in rank order). Builds the
 * global table (identical on every rank; *table_out optional, the caller destroys it), encodes the shard at its global
 * bit offset into d_local and fills *layout. One all-gather (the histograms) and one host round trip (the tree build)
 * sit between the histogram and the encoder. prepare_decode != 0 also flattens the decoder's tables for this table
 * while the encoder runs. */
int mh_sharded_compress(mh_comm* c, const uint8_t* d_in, uint64_t n, int order, uint8_t* d_local, uint64_t local_cap,
                        mh_shard_layout* layout, mh_table** table_out, int prepare_decode, mh_stream_t stream);
/* Collective. Decodes this rank's bit range [bit_base[r], bit_base[r] + n_bits[r]) of the logical stream to d_out.
 * speculative == 0: the layout is exact (codeword boundaries, known contexts): every rank decodes from its known state.
 * speculative != 0: the cuts are treated as arbitrary, as for a stream without an index: neighbours exchange a halo
 * (one all-gather), every rank but the first starts MH_DECODE_WARM_UNIT bits before its range from a guessed state, and
 * the ranks all-gather their seam states; a rank whose warm-up did not reach its predecessor's end state decodes again
 * from exactly that state. The codeword that straddles a cut belongs to the earlier shard. *n_out = symbols this rank
 * wrote, *out_offset (optional) = symbols of the ranks before it. d_local is modified (the halos are spliced in). */
int mh_sharded_decompress(mh_comm* c, const mh_table* t, uint8_t* d_local, uint64_t local_cap, const mh_shard_layout* layout,
                          int speculative, uint8_t* d_out, uint64_t out_capacity, uint64_t* n_out, uint64_t* out_offset,
                          mh_stream_t stream);

/* Ranks of ONE process on shared HOST buffers — what a command-line driver with several GPUs calls in place of
 * compress(FILE*, FILE*) / decompress(FILE*, FILE*) (src/main.cpp:204-212). Only for comms made by mh_comm_create_local:
 * every rank (its own host thread) passes the SAME pointers and sizes; rank r copies its byte range of `in` (compress) or
 * its bit range of the payload plus the warm-up in front of it (decompress) to its GPU, the ranks run the collective calls
 * above (decompress: cuts at arbitrary bits, speculative start, seam handshake), and every rank copies its part of the
 * result to its place in `out` (seam bytes OR-merged). The calls return on every rank when `out` is complete;
 * *out_len = size of the whole result; *table_out (optional) is handed to rank 0 only.
 * mh_sharded_decompress_host answers MH_ERR_INVALID_ARG for a stream too short to cut (use a session on one GPU). */
int mh_sharded_compress_host(mh_comm* c, const uint8_t* in, uint64_t n, int order, uint8_t* out, uint64_t out_capacity,
                             uint64_t* out_len, mh_table** table_out);
int mh_sharded_decompress_host(mh_comm* c, const mh_table* t, const uint8_t* stream, uint64_t stream_len, uint8_t* out,
                               uint64_t out_capacity, uint64_t* out_len);

/* ---------------------------------------------------------------------------------------------------------
 * Synthetic workloads of the benchmark configs; not part of the reference. Byte-identical to
 * the oracle's generators (oracle/mh_oracle.c) so the CPU baseline can be fed the same data.
 * ------------------------------------------------------------------------------------------------------- */
int mh_synth_markov(const uint32_t* trans_counts /* host [65536] */, uint64_t seed, uint64_t seg_bytes,
                    uint64_t first_seg, uint8_t* d_out, uint64_t n, mh_stream_t stream);
int mh_synth_fibonacci(int k_symbols, uint8_t base, uint64_t seed, uint64_t first_index,
                       uint8_t* d_out, uint64_t n, mh_stream_t stream);

/* Launch counter: number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t mh_kernel_launches(void);

/* Per-kernel device timing for bench.py's roofline: while enabled, every launch is bracketed by CUDA events on the
 * stream it is launched on. mh_profile_report synchronises those events and writes a JSON object
 * {"kernel": {"launches": n, "ms": total}, ...} (NUL-terminated) and clears the record. */
int mh_profile_enable(int on);
int mh_profile_report(char* out, size_t cap, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
