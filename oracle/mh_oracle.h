/* oracle/mh_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in plain C, of the reference's Markov-Huffman hot path (jeremy-rifkin/Markov-Huffman-Coding,
 * /root/reference/src). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this; the product (markov-huffman-coding_b200/) never does.
 *
 * Parity status: PINNED. tests/test_oracle.py checks this restatement against
 *   - the golden vectors of SURVEY.md App. C (sizes + SHA-256 of the stock reference's -d outputs for all five
 *     files of /root/reference/test/input in both modes, hex of the tiny cases, edge vectors), committed under
 *     tests/golden/ with the script that generated them from the reference build (tests/golden/make_golden.py);
 *   - the reference itself (oracle/_ref/libmh_ref.so, built from the reference's own sources) on random count
 *     vectors and random inputs whenever oracle/_ref is present.
 *
 * Two deliberate readings of the reference: the table loader reads the LEFT subtree
 * first (the writer and the format comment, src/huffman.cpp:75-81,174-188, define that order; the loader's
 * argument-evaluation order at :170 is unspecified C++), and bit lengths are 64-bit (src/coding.cpp:115,120 use
 * int). Counts and node weights are int32 and wrap exactly as the reference's do (F3).
 */
#ifndef MH_ORACLE_H
#define MH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHO_MAX_NODES 511  /* 256 leaves + 255 internal (the single-symbol hack uses 3) */
#define MHO_CODE_BYTES 32  /* a codeword can be up to 255 bits deep */

typedef struct {
	int16_t left, right;   /* child node indices, -1 for a leaf            (src/tree.h:10-11) */
	uint8_t is_internal;   /*                                               (src/tree.h:12)    */
	uint8_t value;         /*                                               (src/tree.h:13)    */
	int32_t weight;        /* int32, wraps                                  (src/tree.h:14)    */
	int32_t height;        /*                                               (src/tree.h:16)    */
	int32_t depth;         /* assigned by the code-assignment DFS           (src/tree.h:18)    */
} mho_node;

typedef struct {
	int n_nodes;                           /* 0 = empty table (src/huffman.cpp:44-46) */
	int root;
	mho_node nodes[MHO_MAX_NODES];
	int32_t code_len[256];                 /* 0 = symbol has no codeword (src/coding.h:12) */
	uint8_t code_bits[256][MHO_CODE_BYTES];/* MSB-first, unused low bits zero (src/coding.cpp:9-27) */
	int16_t lut[256];                      /* node index per 8-bit window, -1 = null (src/huffman.cpp:12-16,108-121) */
} mho_tree;

typedef struct {
	int markov;          /* get_type(): 1 Markov (256 trees), 0 plain Huffman (1 tree) */
	mho_tree* trees;     /* [256] or [1] */
} mho_table;

/* src/main.cpp:29-39 with the lambdas at :168-170 (order 0) and :176-178 (order 1); prev starts at prev0
 * (' ' in the reference, :32). counts has 256 or 65536 int32 entries and is ADDED to (wrapping). */
void mho_histogram(const uint8_t* in, uint64_t n, uint8_t prev0, int markov, int32_t* counts);

/* src/huffman.cpp:131-164 + src/min_pq.tpp + src/huffman.cpp:97-123. */
void mho_tree_build(mho_tree* t, const int32_t* counts256);

/* huffman_table(int*) / markov_huffman_table(int*): src/huffman.cpp:18-20, src/markov_huffman.cpp:9-13. */
mho_table* mho_table_from_counts(const int32_t* counts, int markov);
/* table-file loader: src/main.cpp:145-161, src/markov_huffman.cpp:15-25, src/huffman.cpp:166-172 (left first).
 * Returns NULL on a truncated / empty file. */
mho_table* mho_table_from_bytes(const uint8_t* buf, size_t n);
void mho_table_free(mho_table* t);

/* table-file writer: src/markov_huffman.cpp:80-88, src/huffman.cpp:174-188. Returns the byte length
 * (bits rounded up, zero padded, src/bitbuffer.cpp:170-180), or -1 if cap is too small. */
long mho_table_write(const mho_table* t, uint8_t* out, size_t cap);

/* i_coding_provider::compress, src/coding.cpp:61-94: out[0] = header, out[1..] = payload.
 * Returns total bytes written (>= 1), or -1 if cap is too small. *dropped (optional) counts symbols whose
 * codeword length is 0 — the reference silently skips them (assert compiled out, src/coding.cpp:72). */
long mho_compress(const mho_table* t, const uint8_t* in, uint64_t n, uint8_t* out, size_t cap, uint64_t* dropped);

/* i_coding_provider::decompress, src/coding.cpp:96-160. Returns the number of bytes decoded, or
 * -1 cap too small, -2 bad signature (:103-106), -3 coder type mismatch (:107-110), -4 a null LUT entry /
 * missing child was reached (the reference would dereference null). */
long mho_decompress(const mho_table* t, const uint8_t* stream, uint64_t stream_len, uint8_t* out, size_t cap);

/* Shard form of the compress loop (test stand-in for one GPU of the sharded path): codes in[0..n) with the byte
 * before in[0] being prev0, writing the payload (no header) so that its first bit lands at bit (bit_base & 7) of
 * out[0]. Returns the number of payload bits, or -1 if cap is too small. out must be zeroed by the caller. */
long long mho_encode_shard(const mho_table* t, const uint8_t* in, uint64_t n, uint8_t prev0, uint64_t bit_base, uint8_t* out, size_t cap);

/* payload bit count a compress() of this input would produce, from counts alone: sum(counts * code_len). */
uint64_t mho_payload_bits(const mho_table* t, const uint8_t* in, uint64_t n);

/* Deterministic synthetic inputs of SURVEY.md §8(d) (bench/test workloads; the GPU generator in the product
 * library must produce the same bytes). See tests/synth.md in DESIGN.md §"Workloads". */
void mho_synth_markov(const uint32_t* trans_counts /*[256*256]*/, uint64_t seed, uint64_t seg_bytes,
                      uint64_t first_seg, uint8_t* out, uint64_t n);
void mho_synth_fibonacci(int k_symbols, uint8_t base, uint64_t seed, uint64_t first_index, uint8_t* out, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
