/* oracle/mh_oracle.c — TEST INFRASTRUCTURE ONLY (see mh_oracle.h for the rules and the parity status: PINNED).
 *
 * A step-by-step CPU restatement of the reference's algorithm for the hot path, written for clarity and
 * bit-exactness, not speed. Every function cites the reference lines it follows (paths under /root/reference).
 */
#include "mh_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------
 * Histogram — src/main.cpp:29-39 (construct_table), lambdas :168-170 (order 0) and :176-178 (order 1).
 * `int` counters: increments wrap like the reference's (done in uint32 arithmetic to stay defined in C).
 * ---------------------------------------------------------------------------------------------------- */
void mho_histogram(const uint8_t* in, uint64_t n, uint8_t prev0, int markov, int32_t* counts) {
	uint32_t* c = (uint32_t*) counts;
	unsigned prev = prev0;                           /* src/main.cpp:32 */
	for(uint64_t i = 0; i < n; i++) {
		if(markov) c[256u * prev + in[i]]++;         /* src/main.cpp:177 */
		else       c[in[i]]++;                       /* src/main.cpp:169 */
		prev = in[i];                                /* src/main.cpp:36  */
	}
}

/* ------------------------------------------------------------------------------------------------------
 * Min priority queue — src/min_pq.tpp. Array binary heap keyed on int32 weight, strict comparisons.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct { int32_t weight; int item; } pq_entry;
typedef struct { pq_entry q[256]; int size; } pq;

static void pq_insert(pq* h, int32_t weight, int item) {      /* src/min_pq.tpp:4-7, swim :29-36 */
	int i = h->size++;
	h->q[i].weight = weight; h->q[i].item = item;
	while(i != 0 && h->q[(i - 1) / 2].weight > h->q[i].weight) {
		pq_entry tmp = h->q[i]; h->q[i] = h->q[(i - 1) / 2]; h->q[(i - 1) / 2] = tmp;
		i = (i - 1) / 2;
	}
}

static int pq_pop_min(pq* h) {                                 /* src/min_pq.tpp:9-15, sink :38-52 */
	pq_entry e = h->q[0];
	h->q[0] = h->q[h->size - 1];
	h->size--;
	int i = 0;
	for(;;) {
		int l = 2 * i + 1, r = 2 * i + 2;
		int target = (r < h->size && h->q[r].weight < h->q[l].weight) ? r : l;   /* ties pick the left child */
		if(target < h->size && h->q[target].weight < h->q[i].weight) {
			pq_entry tmp = h->q[i]; h->q[i] = h->q[target]; h->q[target] = tmp;
			i = target;
		} else break;
	}
	return e.item;
}

/* ------------------------------------------------------------------------------------------------------
 * Codeword record — src/coding.h:9-16, src/coding.cpp:9-27 (push_bit / pop_bit on an MSB-first byte string).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct { int length; uint8_t bits[MHO_CODE_BYTES]; } codeword;

static void cw_push(codeword* w, int b) {                      /* src/coding.cpp:9-16 */
	if(w->length % 8 == 0) w->bits[w->length / 8] = (uint8_t) (b << 7);
	else w->bits[w->length / 8] |= (uint8_t) (b << (8 - w->length % 8 - 1));
	w->length++;
}

static void cw_pop(codeword* w) {                              /* src/coding.cpp:18-27 */
	w->length--;
	if(w->length % 8 == 0) w->bits[w->length / 8] = 0;          /* the byte is dropped (pop_back) */
	else w->bits[w->length / 8] &= (uint8_t) ~(1 << (8 - w->length % 8 - 1));
}

/* ------------------------------------------------------------------------------------------------------
 * Code assignment + 8-bit decode LUT — src/huffman.cpp:91-123. Pre-order DFS, left = 0, right = 1.
 * ---------------------------------------------------------------------------------------------------- */
static void assign_codes(mho_tree* t, int node, codeword* w, int depth) {
	if(node < 0) return;                                        /* :102 */
	mho_node* nd = &t->nodes[node];
	nd->depth = depth;                                          /* :103 */
	if(nd->is_internal) {
		cw_push(w, 0); assign_codes(t, nd->left, w, depth + 1); cw_pop(w);    /* :105-107 */
		cw_push(w, 1); assign_codes(t, nd->right, w, depth + 1); cw_pop(w);   /* :108-110 */
		if(depth == 8) t->lut[w->bits[0]] = (int16_t) node;     /* :111-113 */
	} else {
		t->code_len[nd->value] = w->length;                     /* :115 (a later visit overwrites) */
		memset(t->code_bits[nd->value], 0, MHO_CODE_BYTES);
		memcpy(t->code_bits[nd->value], w->bits, (size_t) (w->length + 7) / 8);
		if(depth <= 8) {                                        /* :116-121 */
			unsigned first = w->bits[0];
			for(int i = 0; i < (1 << (8 - depth)); i++) t->lut[first + i] = (int16_t) node;
		}
	}
}

static void tree_reset(mho_tree* t) {                           /* src/huffman.cpp:12-16 */
	memset(t, 0, sizeof *t);
	t->root = -1;
	for(int i = 0; i < 256; i++) t->lut[i] = -1;
}

static int new_leaf(mho_tree* t, uint8_t v, int32_t w) {        /* src/tree.h:22-23 */
	mho_node* n = &t->nodes[t->n_nodes];
	n->left = n->right = -1; n->is_internal = 0; n->value = v; n->weight = w; n->height = 0; n->depth = -1;
	return t->n_nodes++;
}

static int new_internal(mho_tree* t, int l, int r) {            /* src/tree.h:19-21 */
	mho_node* n = &t->nodes[t->n_nodes];
	n->left = (int16_t) l; n->right = (int16_t) r; n->is_internal = 1; n->value = 0;
	n->weight = (int32_t) ((uint32_t) t->nodes[l].weight + (uint32_t) t->nodes[r].weight);   /* int, wraps */
	n->height = (t->nodes[l].height > t->nodes[r].height ? t->nodes[l].height : t->nodes[r].height) + 1;
	n->depth = -1;
	return t->n_nodes++;
}

static void finish_codes(mho_tree* t) {                         /* src/huffman.cpp:91-95 */
	codeword w; memset(&w, 0, sizeof w);
	assign_codes(t, t->root, &w, 0);
}

/* src/huffman.cpp:131-164 */
void mho_tree_build(mho_tree* t, const int32_t* counts) {
	tree_reset(t);
	pq h; h.size = 0;
	for(int i = 0; i < 256; i++)                                /* :134-138, `if(counts[i])`: non-zero, sign ignored */
		if(counts[i]) pq_insert(&h, counts[i], new_leaf(t, (uint8_t) i, counts[i]));
	if(h.size == 0) return;                                     /* :140-142 */
	while(h.size > 1) {                                         /* :143-151 */
		int a = pq_pop_min(&h), b = pq_pop_min(&h);
		if(t->nodes[a].height > t->nodes[b].height) { int s = a; a = b; b = s; }   /* :147-149 */
		int n = new_internal(t, a, b);
		pq_insert(&h, t->nodes[n].weight, n);                   /* :150: key a->weight + b->weight == node weight */
	}
	t->root = pq_pop_min(&h);                                   /* :152 */
	if(!t->nodes[t->root].is_internal) {                        /* :154-162 single-symbol hack */
		mho_node* r = &t->nodes[t->root];
		int l = new_leaf(t, r->value, r->weight);
		int rr = new_leaf(t, r->value, r->weight);
		r = &t->nodes[t->root];
		r->left = (int16_t) l; r->right = (int16_t) rr; r->height = 1; r->is_internal = 1;
	}
	finish_codes(t);                                            /* :163 */
}

mho_table* mho_table_from_counts(const int32_t* counts, int markov) {
	mho_table* t = (mho_table*) malloc(sizeof *t);
	t->markov = markov ? 1 : 0;
	int ntab = markov ? 256 : 1;
	t->trees = (mho_tree*) malloc(sizeof(mho_tree) * (size_t) ntab);
	for(int i = 0; i < ntab; i++) mho_tree_build(&t->trees[i], counts + 256 * i);   /* src/markov_huffman.cpp:9-13 */
	return t;
}

void mho_table_free(mho_table* t) {
	if(!t) return;
	free(t->trees);
	free(t);
}

/* ------------------------------------------------------------------------------------------------------
 * Bit I/O on memory — the subset of src/bitbuffer.cpp the formats need. MSB-first: stream bit p lives in
 * byte p/8 at bit 7 - p%8 (src/bitbuffer.cpp:12). Writers start from a zeroed buffer (:182-186) and round
 * up to a whole byte on flush (:170-180).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct { uint8_t* buf; size_t cap; uint64_t bit; int overflow; } bitw;
typedef struct { const uint8_t* buf; uint64_t nbits; uint64_t bit; int underflow; } bitr;

static void bw_bit(bitw* w, int b) {                            /* push_bit :9-19 */
	if(w->bit / 8 >= w->cap) { w->overflow = 1; return; }
	if(w->bit % 8 == 0) w->buf[w->bit / 8] = 0;
	w->buf[w->bit / 8] |= (uint8_t) ((b & 1) << (7 - w->bit % 8));
	w->bit++;
}
static void bw_byte(bitw* w, uint8_t v) { for(int i = 7; i >= 0; i--) bw_bit(w, (v >> i) & 1); }   /* push_byte :21-43 */

static int br_bit(bitr* r) {                                    /* pop_bit :82-90 */
	if(r->bit >= r->nbits) { r->underflow = 1; return 0; }
	int b = (r->buf[r->bit / 8] >> (7 - r->bit % 8)) & 1;
	r->bit++;
	return b;
}
static uint8_t br_byte(bitr* r) { uint8_t v = 0; for(int i = 0; i < 8; i++) v = (uint8_t) (v << 1 | br_bit(r)); return v; }  /* pop_byte :92-114 */

/* ------------------------------------------------------------------------------------------------------
 * Table file — writer src/huffman.cpp:174-188 (pre-order: internal -> 0, leaf -> 1 + 8-bit value, left then
 * right) and src/markov_huffman.cpp:80-88 (leading 1, then per prev byte: 0 = empty, 1 + tree).
 * ---------------------------------------------------------------------------------------------------- */
static void write_tree(const mho_tree* t, int node, bitw* w) {
	if(node < 0) return;                                        /* :175-177 */
	const mho_node* n = &t->nodes[node];
	if(n->is_internal) bw_bit(w, 0);                            /* :178-179 */
	else { bw_bit(w, 1); bw_byte(w, n->value); }                /* :180-183 */
	write_tree(t, n->left, w);                                  /* :186 */
	write_tree(t, n->right, w);                                 /* :187 */
}

long mho_table_write(const mho_table* t, uint8_t* out, size_t cap) {
	bitw w = { out, cap, 0, 0 };
	if(t->markov) {
		bw_bit(&w, 1);                                          /* src/markov_huffman.cpp:81 */
		for(int i = 0; i < 256; i++) {
			int nonempty = t->trees[i].n_nodes != 0;
			bw_bit(&w, nonempty);                               /* :83 */
			if(nonempty) write_tree(&t->trees[i], t->trees[i].root, &w);   /* :84-86 */
		}
	} else {
		write_tree(&t->trees[0], t->trees[0].root, &w);         /* src/huffman.cpp:83-85 */
	}
	if(w.overflow) return -1;
	return (long) ((w.bit + 7) / 8);
}

/* loader: src/huffman.cpp:166-172, LEFT subtree first (F1). Leaves get weight 0 (:168). */
static int read_tree(mho_tree* t, bitr* r, int depth) {
	if(r->underflow || t->n_nodes >= MHO_MAX_NODES || depth > 255) { r->underflow = 1; return -1; }
	if(br_bit(r)) {
		uint8_t v = br_byte(r);
		return new_leaf(t, v, 0);
	}
	int self = t->n_nodes++;                                    /* reserve the slot: pre-order numbering */
	int l = read_tree(t, r, depth + 1);
	int rr = read_tree(t, r, depth + 1);
	if(l < 0 || rr < 0) return -1;
	mho_node* n = &t->nodes[self];
	n->left = (int16_t) l; n->right = (int16_t) rr; n->is_internal = 1; n->value = 0; n->weight = 0;
	n->height = (t->nodes[l].height > t->nodes[rr].height ? t->nodes[l].height : t->nodes[rr].height) + 1;
	n->depth = -1;
	return self;
}

mho_table* mho_table_from_bytes(const uint8_t* buf, size_t n) {
	if(n == 0) return NULL;                                     /* the reference would read past EOF */
	bitr r = { buf, (uint64_t) n * 8, 0, 0 };
	int markov = (buf[0] >> 7) & 1;                             /* peek_bit, src/main.cpp:147 */
	mho_table* t = (mho_table*) malloc(sizeof *t);
	t->markov = markov;
	int ntab = markov ? 256 : 1;
	t->trees = (mho_tree*) malloc(sizeof(mho_tree) * (size_t) ntab);
	for(int i = 0; i < ntab; i++) tree_reset(&t->trees[i]);
	if(markov) {
		br_bit(&r);                                             /* src/markov_huffman.cpp:17 */
		for(int i = 0; i < 256; i++)
			if(br_bit(&r)) {                                    /* :20 */
				t->trees[i].root = read_tree(&t->trees[i], &r, 0);
				if(t->trees[i].root >= 0) finish_codes(&t->trees[i]);   /* src/huffman.cpp:22-25 */
			}
	} else {
		t->trees[0].root = read_tree(&t->trees[0], &r, 0);
		if(t->trees[0].root >= 0) finish_codes(&t->trees[0]);
	}
	if(r.underflow) { mho_table_free(t); return NULL; }
	return t;
}

/* ------------------------------------------------------------------------------------------------------
 * compress — src/coding.cpp:61-94 with push_encoding_descriptor (src/bitbuffer.cpp:45-73) reduced to what it
 * does to the stream: append the codeword's bits MSB-first.
 * ---------------------------------------------------------------------------------------------------- */
long mho_compress(const mho_table* t, const uint8_t* in, uint64_t n, uint8_t* out, size_t cap, uint64_t* dropped) {
	if(cap < 1) return -1;
	memset(out, 0, cap < 1 + n ? cap : (size_t) (1 + n));       /* cheap pre-zero of the common case */
	bitw w = { out + 1, cap - 1, 0, 0 };
	uint64_t miss = 0;
	unsigned prev = ' ';                                        /* :67 */
	for(uint64_t i = 0; i < n; i++) {
		const mho_tree* tr = &t->trees[t->markov ? prev : 0];  /* get_encoding: src/markov_huffman.cpp:52-54, src/huffman.cpp:71-73 */
		unsigned c = in[i];
		int len = tr->code_len[c];
		if(len == 0) miss++;                                    /* assert compiled out (:72): nothing is emitted */
		prev = c;                                               /* :74 */
		const uint8_t* bits = tr->code_bits[c];
		for(int b = 0; b < len; b++) bw_bit(&w, (bits[b / 8] >> (7 - b % 8)) & 1);   /* :76 */
	}
	if(w.overflow) return -1;
	int bi = (int) (w.bit % 8);                                 /* get_bi, :85 */
	out[0] = (uint8_t) (0x30 | ((~t->markov & 1) << 3) | ((8 - bi) % 8));   /* :88 */
	if(dropped) *dropped = miss;
	return 1 + (long) ((w.bit + 7) / 8);
}

long long mho_encode_shard(const mho_table* t, const uint8_t* in, uint64_t n, uint8_t prev0, uint64_t bit_base, uint8_t* out, size_t cap) {
	uint64_t pos = bit_base & 7;                                /* the shard's first bit inside out[0] */
	const uint64_t start = pos;
	unsigned prev = prev0;
	for(uint64_t i = 0; i < n; i++) {
		const mho_tree* tr = &t->trees[t->markov ? prev : 0];
		unsigned c = in[i];
		int len = tr->code_len[c];
		const uint8_t* bits = tr->code_bits[c];
		for(int b = 0; b < len; b++, pos++) {
			if(pos / 8 >= cap) return -1;
			out[pos / 8] |= (uint8_t) (((bits[b / 8] >> (7 - b % 8)) & 1) << (7 - pos % 8));
		}
		prev = c;
	}
	return (long long) (pos - start);
}

uint64_t mho_payload_bits(const mho_table* t, const uint8_t* in, uint64_t n) {
	uint64_t bits = 0; unsigned prev = ' ';
	for(uint64_t i = 0; i < n; i++) { bits += (uint64_t) t->trees[t->markov ? prev : 0].code_len[in[i]]; prev = in[i]; }
	return bits;
}

/* ------------------------------------------------------------------------------------------------------
 * decompress — src/coding.cpp:96-160. The window `w` is the next 8 stream bits, zero padded past the end of
 * the FILE (pop_rest / try_pop_bit, src/bitbuffer.cpp:116-140), which is why the reader below pads with
 * zeros beyond stream_len and never beyond `length` only.
 * ---------------------------------------------------------------------------------------------------- */
static int stream_bit(const uint8_t* payload, uint64_t payload_bits_total, uint64_t p) {
	if(p >= payload_bits_total) return 0;
	return (payload[p / 8] >> (7 - p % 8)) & 1;
}

long mho_decompress(const mho_table* t, const uint8_t* stream, uint64_t stream_len, uint8_t* out, size_t cap) {
	if(stream_len < 1) return -2;
	uint8_t header = stream[0];                                 /* :100 */
	if((header & 0xF0) != 0x30) return -2;                      /* :103-106 */
	if(((~(header & (1 << 3)) >> 3) & 1) != t->markov) return -3;   /* :107-110 */
	int remainder = header & 7;                                 /* :111 */
	int64_t length = ((int64_t) stream_len - 1) * 8 - remainder;   /* :115, 64-bit (F2) */
	const uint8_t* payload = stream + 1;
	uint64_t file_bits = (stream_len - 1) * 8;
	unsigned prev = ' ';                                        /* :118 */
	int64_t bi = 0;                                             /* :120 */
	size_t n_out = 0;
	while(bi < length) {                                        /* :124 */
		unsigned w = 0;
		for(int k = 0; k < 8; k++) w = w << 1 | (unsigned) stream_bit(payload, file_bits, (uint64_t) bi + k);   /* :125 */
		const mho_tree* tr = &t->trees[t->markov ? prev : 0];  /* decoding_lookup: src/markov_huffman.cpp:56-58 */
		int node = tr->lut[w];                                  /* src/huffman.cpp:87-89 */
		if(node < 0) return -4;
		if(tr->nodes[node].is_internal) {                       /* :129-149 */
			bi += 8;
			while(tr->nodes[node].is_internal) {
				int bit = stream_bit(payload, file_bits, (uint64_t) bi);
				bi++;
				node = bit ? tr->nodes[node].right : tr->nodes[node].left;
				if(node < 0) return -4;
			}
		} else {
			bi += tr->nodes[node].depth;                        /* :155 */
		}
		if(n_out >= cap) return -1;
		out[n_out++] = tr->nodes[node].value;                   /* :139 / :151 */
		prev = tr->nodes[node].value;                           /* :140 / :152 */
	}
	return (long) n_out;
}

/* ------------------------------------------------------------------------------------------------------
 * Synthetic workloads. Not part of the reference; shared definition with the product's
 * GPU generator (csrc/mh_synth.cu), which tests compare byte for byte against these.
 *   rnd(i)   = splitmix64 finaliser of (seed + (i + 1) * 0x9E3779B97F4A7C15), i = global byte index
 *   target   = (hi32(rnd) * row_total) >> 32                        in [0, row_total)
 *   symbol   = smallest c with cumulative_count[c] > target
 * Markov text: segment s covers bytes [s*seg_bytes, (s+1)*seg_bytes); every segment restarts from context ' '.
 * A context with no successors falls back to the ' ' row.
 * ---------------------------------------------------------------------------------------------------- */
static uint64_t mix64(uint64_t z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
	z ^= z >> 27; z *= 0x94D049BB133111EBull;
	z ^= z >> 31;
	return z;
}
static uint64_t rnd_at(uint64_t seed, uint64_t i) { return mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ull); }

void mho_synth_markov(const uint32_t* tc, uint64_t seed, uint64_t seg_bytes, uint64_t first_seg, uint8_t* out, uint64_t n) {
	static uint64_t cum[256][256];
	uint64_t total[256];
	for(int p = 0; p < 256; p++) {
		uint64_t s = 0;
		for(int c = 0; c < 256; c++) { s += tc[256 * p + c]; cum[p][c] = s; }
		total[p] = s;
	}
	for(uint64_t off = 0; off < n; off += seg_bytes) {
		uint64_t seg = first_seg + off / seg_bytes;
		unsigned prev = ' ';
		uint64_t len = n - off < seg_bytes ? n - off : seg_bytes;
		for(uint64_t j = 0; j < len; j++) {
			uint64_t gi = seg * seg_bytes + j;
			unsigned row = total[prev] ? prev : ' ';
			unsigned sym = ' ';
			if(total[row]) {
				uint64_t target = ((rnd_at(seed, gi) >> 32) * total[row]) >> 32;
				unsigned c = 0;
				while(cum[row][c] <= target) c++;
				sym = c;
			}
			out[off + j] = (uint8_t) sym;
			prev = sym;
		}
	}
}

void mho_synth_fibonacci(int k, uint8_t base, uint64_t seed, uint64_t first_index, uint8_t* out, uint64_t n) {
	uint64_t cum[64], a = 1, b = 1, s = 0;                      /* weight of symbol j = Fib(j+1): 1,1,2,3,5,... */
	if(k > 64) k = 64;
	for(int j = 0; j < k; j++) { s += a; cum[j] = s; uint64_t nx = a + b; a = b; b = nx; }
	for(uint64_t i = 0; i < n; i++) {
		uint64_t target = ((rnd_at(seed, first_index + i) >> 32) * s) >> 32;
		int j = 0;
		while(cum[j] <= target) j++;
		out[i] = (uint8_t) (base + j);
	}
}
