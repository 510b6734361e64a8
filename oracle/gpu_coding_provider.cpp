// oracle/gpu_coding_provider.cpp — the binding of INTEGRATION.md §2, compiled for real.
//
// TEST INFRASTRUCTURE (built by `make -C oracle ref_gpu` into oracle/_ref/markovhuffman_gpu): this file takes the
// place of the reference's src/coding.cpp when the reference's OWN main.cpp, huffman.cpp, markov_huffman.cpp,
// tree.cpp, bitbuffer.cpp and utils.cpp are compiled from /root/reference/src, unmodified (huffman.cpp with the
// F1 sequencing fix that the patched oracle build uses). It defines exactly what src/coding.cpp defines:
// the three encoding_descriptor helpers (src/coding.cpp:9-33) and the two hot loops of the abstract provider,
// i_coding_provider::compress / ::decompress (src/coding.h:26-27, src/coding.cpp:61-160) — here bound to libmh_gpu.so's
// C ABI (include/mh_gpu.h). The reference's CLI, histogram loop, tree builder, table loader / writer and -g printers
// all stay the reference's; only the two loops run on the GPU. tests/test_reference_dropin.py runs the reference's
// own test flow (test/main.py:68-105) through the resulting binary.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "coding.h"   // the reference's header, from /root/reference/src
#include "utils.h"

#include "mh_gpu.h"

// ---- encoding_descriptor (src/coding.h:9-16): a codeword as MSB-first bytes plus its bit length --------------------
void encoding_descriptor::push_bit(int b) {
	const int at = length & 7;
	if(at == 0) encoding.push_back(static_cast<unsigned char>(b << 7));
	else encoding[length >> 3] |= static_cast<unsigned char>(b << (7 - at));
	++length;
}

void encoding_descriptor::pop_bit() {
	--length;
	const size_t need = static_cast<size_t>((length + 7) >> 3);
	if(need < encoding.size()) encoding.pop_back();
	else encoding[length >> 3] &= static_cast<unsigned char>(~(0x80u >> (length & 7)));
}

void encoding_descriptor::print() {
	for(int k = 0; k < length; ++k) putchar('0' + ((encoding[k >> 3] >> (7 - (k & 7))) & 1));
}

namespace {

std::vector<unsigned char> read_all(FILE* f) {
	std::vector<unsigned char> bytes;
	unsigned char chunk[1 << 16];
	size_t got;
	while((got = fread(chunk, 1, sizeof chunk, f)) > 0) bytes.insert(bytes.end(), chunk, chunk + got);
	return bytes;
}

[[noreturn]] void fail(const char* what, int rc) {
	if(rc == MH_ERR_BAD_HEADER) eprintf("Error while decoding file: Input appears corrupt.\n");                          // src/coding.cpp:104
	else if(rc == MH_ERR_TYPE_MISMATCH) eprintf("Error: File encoding method does not match provided encoding table.\n");   // src/coding.cpp:108
	else eprintf("Error: %s: %s (%s)\n", what, mh_status_string(rc), mh_last_error());
	exit(1);
}

// The provider's trees, as libmh_gpu's table: through the provider's own serialiser (write_coding_tree, the table-file
// format is the contract) into a scratch file, and back in with mh_table_from_bytes. The bitbuffer closes the file it
// is given (src/bitbuffer.h:35-40), so the scratch file is reopened by name.
mh_table* table_of(i_coding_provider& provider) {
	char path[] = "/tmp/mh_table_XXXXXX";
	const int fd = mkstemp(path);
	if(fd < 0) { eprintf("Error: cannot create a scratch file; %s.\n", strerror(errno)); exit(1); }
	{
		bitbuffer sink(fdopen(fd, "wb"), bitbuffer::write);
		provider.write_coding_tree(sink);
	}   // flushed and closed here
	FILE* back = fopen(path, "rb");
	std::vector<unsigned char> image = back ? read_all(back) : std::vector<unsigned char>();
	if(back) fclose(back);
	unlink(path);
	mh_table* t = nullptr;
	const int rc = mh_table_from_bytes(image.data(), image.size(), &t);
	if(rc != MH_OK) fail("loading the provider's table", rc);
	return t;
}

}  // namespace

// replaces the loop at src/coding.cpp:61-94: header byte + MSB-first payload, same bytes
void i_coding_provider::compress(FILE* input_fd, FILE* output_fd) {
	mh_table* t = table_of(*this);
	std::vector<unsigned char> in = read_all(input_fd);
	fclose(input_fd);
	mh_session* s = nullptr;
	int rc = mh_session_create(0, in.size() + 64, &s);
	if(rc != MH_OK) fail("creating the GPU session", rc);
	std::vector<unsigned char> out(in.size() + in.size() / 8 + 4200);
	uint64_t out_len = 0, dropped = 0;
	rc = mh_session_compress_with_table(s, t, in.data(), in.size(), out.data(), out.size(), &out_len, &dropped);
	if(rc == MH_ERR_CAPACITY && out_len > out.size()) {   // a foreign table that expands the input: the call reported the size it needs
		out.resize(out_len);
		rc = mh_session_compress_with_table(s, t, in.data(), in.size(), out.data(), out.size(), &out_len, &dropped);
	}
	if(rc != MH_OK) fail("compressing", rc);
	write_buffer(out.data(), 1, out_len, output_fd);
	mh_session_destroy(s);
	mh_table_destroy(t);
}

// replaces the loop at src/coding.cpp:96-160 (header checks included: the library returns the reference's two errors)
void i_coding_provider::decompress(FILE* input_fd, FILE* output_fd) {
	mh_table* t = table_of(*this);
	std::vector<unsigned char> stream = read_all(input_fd);
	fclose(input_fd);
	mh_session* s = nullptr;
	int rc = mh_session_create_sized(0, stream.size() * 3 + 4096, stream.size() + 64, &s);
	if(rc != MH_OK) fail("creating the GPU session", rc);
	uint64_t n = 0;
	rc = mh_session_decompress(s, t, stream.data(), stream.size(), nullptr, 0, &n);   // decode on the device, learn the size
	if(rc != MH_OK && rc != MH_ERR_CORRUPT_STREAM) fail("extracting", rc);
	std::vector<unsigned char> out(n ? n : 1);
	rc = mh_session_fetch(s, out.data(), out.size(), &n);
	if(rc == MH_ERR_WORKSPACE) rc = mh_session_decompress(s, t, stream.data(), stream.size(), out.data(), out.size(), &n);   // it was decoded in chunks: again, into the buffer
	if(rc != MH_OK && rc != MH_ERR_CORRUPT_STREAM) fail("extracting", rc);
	write_buffer(out.data(), 1, n, output_fd);
	mh_session_destroy(s);
	mh_table_destroy(t);
}
