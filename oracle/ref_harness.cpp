// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" shim (our own code) over the reference's UNMODIFIED table classes, linked against
// objects compiled straight from /root/reference/src by oracle/Makefile (decode-side objects carry the
// two on-the-fly fixes F1/F2 described there). It lets the tests drive the real reference with arbitrary
// count vectors, in memory, without the CLI:
//   huffman_table(int*)            /root/reference/src/huffman.h:14
//   markov_huffman_table(int*)     /root/reference/src/markov_huffman.h:12
//   write_coding_tree(bitbuffer&)  /root/reference/src/coding.h:23
//   compress / decompress          /root/reference/src/coding.h:26-27
//   get_encoding / decoding_lookup /root/reference/src/huffman.h:26-27 (public in the concrete classes)
// Nothing here is shipped or measured as product code.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <string>
#include <vector>

#include "bitbuffer.h"
#include "coding.h"
#include "huffman.h"
#include "markov_huffman.h"

namespace {

std::string tmp_path(const char* tag) {
	const char* dir = access("/dev/shm", W_OK) == 0 ? "/dev/shm" : "/tmp";
	std::string p = std::string(dir) + "/mhref_" + tag + "_XXXXXX";
	int fd = mkstemp(&p[0]);
	if(fd >= 0) close(fd);
	return p;
}

void spill(const std::string& path, const unsigned char* data, size_t n) {
	FILE* f = fopen(path.c_str(), "wb");
	if(n) fwrite(data, 1, n, f);
	fclose(f);
}

long slurp(const std::string& path, unsigned char* out, size_t cap) {
	FILE* f = fopen(path.c_str(), "rb");
	if(!f) return -1;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	if((size_t) n > cap) { fclose(f); return -2 - n; }
	if(n) { size_t got = fread(out, 1, n, f); (void) got; }
	fclose(f);
	return n;
}

i_coding_provider* provider_from_counts(const int* counts, int markov) {
	std::vector<int> copy(counts, counts + (markov ? 65536 : 256));
	if(markov) return new markov_huffman_table(copy.data());
	return new huffman_table(copy.data());
}

// The table file's first bit tells the kind (src/main.cpp:147-161). A 0-byte file is an empty -h table,
// which the reference's loader cannot read (it would pop a bit from an empty file); the harness refuses it.
i_coding_provider* provider_from_table(const std::string& path, long size, int* markov) {
	if(size <= 0) return nullptr;
	FILE* f = fopen(path.c_str(), "rb");
	bitbuffer* bb = new bitbuffer(f, bitbuffer::read);
	i_coding_provider* p;
	if(bb->peek_bit()) { *markov = 1; p = new markov_huffman_table(*bb); }
	else               { *markov = 0; p = new huffman_table(*bb); }
	delete bb; // closes f
	return p;
}

} // namespace

extern "C" {

// Build the provider from counts and serialise the table exactly as `-d` does (src/main.cpp:193-202).
// Returns the table length, or <0 on error / insufficient capacity.
long ref_table_from_counts(const int* counts, int markov, unsigned char* table, size_t cap) {
	i_coding_provider* p = provider_from_counts(counts, markov);
	std::string path = tmp_path("tab");
	{
		bitbuffer bb(fopen(path.c_str(), "wb"), bitbuffer::write);
		// an empty -h table has a null tree: write_coding_tree_traversal returns immediately (huffman.cpp:175)
		p->write_coding_tree(bb);
	}
	long n = slurp(path, table, cap);
	unlink(path.c_str());
	delete p;
	return n;
}

// Codewords as the reference derives them (src/huffman.cpp:97-123). lens[t*256+c] = bit length (0 = none),
// bits[(t*256+c)*32 ..] = MSB-first bytes. ntab = 1 (-h) or 256 (Markov).
int ref_codes_from_counts(const int* counts, int markov, int* lens, unsigned char* bits) {
	int ntab = markov ? 256 : 1;
	memset(lens, 0, sizeof(int) * ntab * 256);
	memset(bits, 0, (size_t) ntab * 256 * 32);
	for(int t = 0; t < ntab; t++) {
		std::vector<int> copy(counts + 256 * t, counts + 256 * t + 256);
		huffman_table h(copy.data());
		for(int c = 0; c < 256; c++) {
			encoding_descriptor& e = h.get_encoding(0, (unsigned char) c);
			lens[t * 256 + c] = e.length;
			for(size_t j = 0; j < e.encoding.size() && j < 32; j++) bits[((size_t) t * 256 + c) * 32 + j] = e.encoding[j];
		}
	}
	return 0;
}

// The 8-bit decode LUT (src/huffman.cpp:108-121): kind[t*256+w] = 0 null, 1 leaf, 2 internal-at-depth-8;
// value/depth as stored in the node.
int ref_lut_from_counts(const int* counts, int markov, unsigned char* kind, unsigned char* value, int* depth) {
	int ntab = markov ? 256 : 1;
	for(int t = 0; t < ntab; t++) {
		std::vector<int> copy(counts + 256 * t, counts + 256 * t + 256);
		huffman_table h(copy.data());
		for(int w = 0; w < 256; w++) {
			const tree_node* n = h.decoding_lookup(0, (unsigned char) w);
			int k = t * 256 + w;
			kind[k] = n == nullptr ? 0 : (n->is_internal ? 2 : 1);
			value[k] = n ? n->value : 0;
			depth[k] = n ? n->depth : 0;
		}
	}
	return 0;
}

// compress() with a provider built from counts (the `-d` path, src/main.cpp:164-183 then :211).
long ref_compress_counts(const int* counts, int markov, const unsigned char* in, size_t n, unsigned char* out, size_t cap) {
	i_coding_provider* p = provider_from_counts(counts, markov);
	std::string pin = tmp_path("in"), pout = tmp_path("out");
	spill(pin, in, n);
	p->compress(fopen(pin.c_str(), "rb"), fopen(pout.c_str(), "wb")); // both FILE*s are consumed
	long r = slurp(pout, out, cap);
	unlink(pin.c_str()); unlink(pout.c_str());
	delete p;
	return r;
}

// compress() with a provider loaded from a table file (the `-e` path, src/main.cpp:137-162 then :211).
long ref_compress_table(const unsigned char* table, size_t tn, const unsigned char* in, size_t n, unsigned char* out, size_t cap) {
	std::string ptab = tmp_path("tab");
	spill(ptab, table, tn);
	int markov = 0;
	i_coding_provider* p = provider_from_table(ptab, (long) tn, &markov);
	unlink(ptab.c_str());
	if(!p) return -1;
	std::string pin = tmp_path("in"), pout = tmp_path("out");
	spill(pin, in, n);
	p->compress(fopen(pin.c_str(), "rb"), fopen(pout.c_str(), "wb"));
	long r = slurp(pout, out, cap);
	unlink(pin.c_str()); unlink(pout.c_str());
	delete p;
	return r;
}

// decompress() with a provider loaded from a table file (the `-x -e` path, src/main.cpp:204-207).
long ref_decompress_table(const unsigned char* table, size_t tn, const unsigned char* stream, size_t sn, unsigned char* out, size_t cap) {
	std::string ptab = tmp_path("tab");
	spill(ptab, table, tn);
	int markov = 0;
	i_coding_provider* p = provider_from_table(ptab, (long) tn, &markov);
	unlink(ptab.c_str());
	if(!p) return -1;
	std::string pin = tmp_path("in"), pout = tmp_path("out");
	spill(pin, stream, sn);
	p->decompress(fopen(pin.c_str(), "rb"), fopen(pout.c_str(), "wb"));
	long r = slurp(pout, out, cap);
	unlink(pin.c_str()); unlink(pout.c_str());
	delete p;
	return r;
}

} // extern "C"
