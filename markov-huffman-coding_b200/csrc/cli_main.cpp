// cli_main.cpp — `markovhuffman`, the C++ host driver behind the reference's command line
//   markov-huffman <input> [-o output] [-h] [-e encoding_file] [-d output_encoding_file] [-g] [-x]
// (reference src/main.cpp:17-27, :41-217). The flag grammar, validation errors, exit codes and stderr progress
// lines follow the reference; all coding work goes through the C ABI of libmh_gpu.so (include/mh_gpu.h) and runs
// on the GPU. There is no CPU path: without a usable device the tool reports the error and exits 1.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../include/mh_gpu.h"

#define eprintf(...) fprintf(stderr, __VA_ARGS__)

namespace {

void usage() {   // src/main.cpp:17-27
	eprintf("markov-huffman <input> [-o output] [options]\n");
	eprintf("\t-o output_file\n");
	eprintf("\t-h use simple huffman coding\n");
	eprintf("\n");
	eprintf("\t-e encoding_file\n");
	eprintf("\t-d output_encoding_file\n");
	eprintf("\n");
	eprintf("\t-g print huffman trees and tables\n");
	eprintf("\t-x extract\n");
}

// src/utils.cpp:44-56: only warns, and without a trailing newline
void probe_access(const char* path, bool write) {
	if(access(path, write ? W_OK : R_OK) == -1)
		eprintf("Error: Unable to open \"%s\" for %s; %s.", path, write ? "writing" : "reading", strerror(errno));
}

// A host buffer in page-locked memory (mh_pinned_alloc), so that the session's copies overlap with its kernels; plain
// memory when pinning fails (the copies are then staged by the driver).
struct HostBuf {
	uint8_t* p = nullptr;
	size_t n = 0, cap = 0;
	bool pinned = false;
	HostBuf() = default;
	HostBuf(const HostBuf&) = delete;
	HostBuf& operator=(const HostBuf&) = delete;
	~HostBuf() { release(); }
	void release() {
		if(pinned) mh_pinned_free(p);
		else free(p);
		p = nullptr;
		n = cap = 0;
	}
	bool reserve(size_t want) {   // contents are kept
		if(want <= cap) return true;
		uint8_t* q = static_cast<uint8_t*>(mh_pinned_alloc(want));
		const bool pin = q != nullptr;
		if(!q) q = static_cast<uint8_t*>(malloc(want));
		if(!q) return false;
		if(n) memcpy(q, p, n);
		const size_t keep = n;
		release();
		p = q; n = keep; cap = want; pinned = pin;
		return true;
	}
	bool resize(size_t want) {
		if(!reserve(want ? want : 1)) return false;
		n = want;
		return true;
	}
};

// the whole file, read straight into the pinned buffer (sized from the file when it is seekable)
bool slurp(FILE* f, HostBuf& out) {
	size_t hint = 1 << 16;
	if(fseek(f, 0, SEEK_END) == 0) {
		const long end = ftell(f);
		if(end > 0) hint = size_t(end) + 1;
		rewind(f);
	}
	if(!out.reserve(hint)) return false;
	size_t got;
	while((got = fread(out.p + out.n, 1, out.cap - out.n, f)) > 0) {
		out.n += got;
		if(out.n == out.cap && !out.reserve(out.cap * 2)) return false;
	}
	return !ferror(f);
}

[[noreturn]] void die_status(const char* what, int rc) {
	if(rc == MH_ERR_BAD_HEADER) eprintf("Error while decoding file: Input appears corrupt.\n");                        // src/coding.cpp:104
	else if(rc == MH_ERR_TYPE_MISMATCH) eprintf("Error: File encoding method does not match provided encoding table.\n");   // src/coding.cpp:108
	else if(rc == MH_ERR_CUDA || rc == MH_ERR_NO_DEVICE) eprintf("Error: %s: %s (%s). This build has no CPU path.\n", what, mh_status_string(rc), mh_last_error());
	else eprintf("Error: %s: %s.\n", what, mh_status_string(rc));
	exit(1);
}

const char* shown(const char* s) { return s ? s : "(null)"; }   // what glibc prints for a null %s

// How many GPUs the file is cut over: MH_CLI_GPUS when set, else every device of the box for inputs of at
// least 256 MiB, else one. The shards have to fit their devices; otherwise the single-GPU session streams the file.
int gpus_for(uint64_t file_bytes, uint64_t per_byte_footprint) {
	int ndev = mh_device_count();
	if(ndev > MH_MAX_SHARDS) ndev = MH_MAX_SHARDS;
	int want = file_bytes >= (256ull << 20) ? ndev : 1;
	if(const char* env = getenv("MH_CLI_GPUS")) want = atoi(env);
	if(want < 2) return 1;
	uint64_t free_b = 0, total_b = 0;
	if(mh_device_memory(0, &free_b, &total_b) != MH_OK) return 1;
	if(file_bytes / uint64_t(want) * per_byte_footprint + (64ull << 20) > free_b / 2) return 1;
	return want;
}

// fn(rank, comm) on one host thread per GPU (or per rank: MH_CLI_GPUS may exceed the device count, the ranks then share
// devices through the library's in-process transport); returns the first failure, MH_OK otherwise
template <typename F>
int on_every_rank(int world, F fn) {
	std::vector<mh_comm*> comms(size_t(world), nullptr);
	int rc = mh_comm_create_local(world, nullptr, 1, comms.data());
	if(rc != MH_OK) return rc;
	std::vector<int> rcs(size_t(world), MH_OK);
	std::vector<std::thread> threads;
	for(int r = 0; r < world; ++r) threads.emplace_back([&, r] { rcs[size_t(r)] = fn(r, comms[size_t(r)]); });
	for(auto& t : threads) t.join();
	for(mh_comm* c : comms) mh_comm_destroy(c);
	for(int v : rcs)
		if(v != MH_OK) return v;
	return MH_OK;
}

}  // namespace

int main(int argc, char* argv[]) {
	if(argc < 2) {
		usage();
		return 1;
	}
	bool extract = false, debug = false, simple_huffman = false;
	const char* input = nullptr;
	const char* output = nullptr;
	const char* encoding_input = nullptr;
	const char* encoding_output = nullptr;
	// src/main.cpp:55-101: short flags may be combined ("-xh"); every value flag inside one argument consumes the
	// next unused argv entry in order; a bare "-" is a no-op; unknown flags warn and continue.
	for(int i = 1; i < argc; i++) {
		if(argv[i][0] != '-') {
			if(!input) input = argv[i];
			else eprintf("Warning: Unexpected positional argument %s.\n", argv[i]);
			continue;
		}
		int taken = 0;
		auto value = [&]() -> const char* {
			const int at = i + taken + 1;
			++taken;
			return at <= argc ? argv[at] : nullptr;   // argv[argc] is the null terminator, as the reference would read
		};
		for(const char* f = argv[i] + 1; *f; ++f) {
			switch(*f) {
				case 'o':
					if(i + 1 < argc) output = value();
					else eprintf("Error: Expected output file following -o.\n");
					break;
				case 'e':
					if(i + 1 < argc) encoding_input = value();
					else eprintf("Error: Expected encoding file following -e.\n");
					break;
				case 'd':
					if(i + 1 < argc) encoding_output = value();
					else eprintf("Error: Expected encoding output file following -d.\n");
					break;
				case 'x': extract = true; break;
				case 'h': simple_huffman = true; break;
				case 'g': debug = true; break;
				default: eprintf("Warning: Unknown option %c.\n", *f);
			}
		}
		i += taken;
	}

	// src/main.cpp:104-115
	if(!input) {
		eprintf("Error: Must provide input file.\n");
		exit(1);
	}
	if(encoding_input && encoding_output) {
		eprintf("Error: Don't provide an encoding input and an encoding output. Just use cp.\n");
		exit(1);
	}
	if(extract && !encoding_input) {
		eprintf("Error: Must provide encoding file input while in decompress mode.\n");
		exit(1);
	}
	// src/main.cpp:118-121
	probe_access(input, false);
	if(output) probe_access(output, true);
	if(encoding_input) probe_access(encoding_input, false);
	if(encoding_output) probe_access(encoding_output, false);

	FILE* input_fd = fopen(input, "rb");
	if(!input_fd) {
		eprintf("Error while opening input; %s.\n", strerror(errno));
		exit(1);
	}
	FILE* output_fd = output ? fopen(output, "wb") : stdout;   // opened early to catch errors (src/main.cpp:129-134)
	if(!output_fd) {
		eprintf("Error while opening output; %s.\n", strerror(errno));
		exit(1);
	}

	HostBuf in_bytes;
	if(!slurp(input_fd, in_bytes)) {
		eprintf("Error occurred while reading file.\n");   // src/utils.cpp:62-65
		exit(1);
	}
	fclose(input_fd);

	const int order = simple_huffman ? MH_ORDER_HUFFMAN : MH_ORDER_MARKOV;
	mh_table* table = nullptr;
	mh_session* session = nullptr;
	HostBuf result;

	if(encoding_input) {
		eprintf("Loading encoding table from file...\n");
		FILE* tf = fopen(encoding_input, "rb");
		if(!tf) {
			eprintf("Error while opening encoding input; %s.\n", strerror(errno));
			exit(1);
		}
		HostBuf tbytes;
		if(!slurp(tf, tbytes)) {
			eprintf("Error occurred while reading file.\n");
			exit(1);
		}
		fclose(tf);
		if(tbytes.n == 0) {   // the reference peeks a bit of an empty buffer here (undefined); refuse instead
			eprintf("Error: Encoding table file is empty.\n");
			exit(1);
		}
		const bool file_is_markov = (tbytes.p[0] & 0x80) != 0;   // src/main.cpp:147-154
		if(file_is_markov != !simple_huffman) {
			eprintf("Error: Incorrect encoding table provided for current operation; expected %s, found %s.\n",
			        simple_huffman ? "simple Huffman" : "Markov-Huffman", file_is_markov ? "Markov-Huffman" : "simple Huffman");
			exit(1);
		}
		int rc = mh_table_from_bytes(tbytes.p, tbytes.n, &table);
		if(rc != MH_OK) die_status("loading the encoding table", rc);
	}

	// Device buffers. Compressing: the input bounds both sides. Extracting: the stream bounds the compressed side; the
	// decoded size is not stored in the stream, so the session grows its uncompressed-side buffer after the count pass.
	// A file that does not fit the GPU goes through the session in chunks (histogram counts add up, every chunk is coded
	// at its global bit offset / decoded from the exact state its predecessor ended in), so the buffers are capped at a
	// fifth of the free device memory; MH_CLI_MAX_BYTES lowers the cap (tests).
	uint64_t cap = ~uint64_t(0), free_b = 0, total_b = 0;
	if(mh_device_memory(0, &free_b, &total_b) == MH_OK && free_b / 5 > (1u << 20)) cap = free_b / 5;
	if(const char* env = getenv("MH_CLI_MAX_BYTES")) {
		const unsigned long long v = strtoull(env, nullptr, 10);
		if(v >= 4096) cap = v;
	}
	auto capped = [&](uint64_t want) { return want < cap ? want : cap; };
	int rc = MH_OK;
	auto open_session = [&]() {
		if(session) return;
		rc = extract ? mh_session_create_sized(0, capped(in_bytes.n * 3 + 4096), capped(in_bytes.n + 64), &session)
		             : mh_session_create(0, capped(in_bytes.n + 64), &session);
		if(rc != MH_OK) die_status("creating the GPU session", rc);
	};
	auto alloc_fail = [&]() { eprintf("Error: out of host memory.\n"); exit(1); };

	bool built_here = false;
	if(!encoding_input) {
		eprintf(simple_huffman ? "Building simple Huffman encoding table from input...\n"
		                       : "Building Markov-Huffman encoding table from input...\n");
		built_here = true;
	}

	if(built_here) {
		// histogram -> trees -> encode; the table comes back for -g / -d
		if(!result.resize(in_bytes.n + in_bytes.n / 8 + 4200)) alloc_fail();
		uint64_t out_len = 0;
		const int gpus = gpus_for(in_bytes.n, 3);
		bool done = false;
		if(gpus > 1) {   // one stream over several GPUs: byte-range shards, one host thread per GPU
			std::vector<uint64_t> lens(size_t(gpus), 0);
			rc = on_every_rank(gpus, [&](int r, mh_comm* c) {
				return mh_sharded_compress_host(c, in_bytes.p, in_bytes.n, order, result.p, result.n, &lens[size_t(r)], r == 0 ? &table : nullptr);
			});
			if(rc == MH_OK) { out_len = lens[0]; done = true; }
			else if(table) { mh_table_destroy(table); table = nullptr; }
		}
		if(!done) {
			open_session();
			rc = mh_session_compress(session, in_bytes.p, in_bytes.n, order, result.p, result.n, &out_len, &table);
			if(rc != MH_OK) die_status("compressing", rc);
		}
		result.n = out_len;
	}

	if(debug) {   // src/main.cpp:187-190
		size_t n = 0;
		mh_table_debug_dump(table, nullptr, 0, &n);
		std::string text(n, '\0');
		if(n && mh_table_debug_dump(table, &text[0], n, &n) == MH_OK) fwrite(text.data(), 1, n, stdout);
	}

	if(encoding_output) {   // src/main.cpp:193-202
		FILE* ef = fopen(encoding_output, "wb");
		eprintf("Writing encoding table to %s...\n", encoding_output);
		if(!ef) {
			eprintf("Error while opening encoding file output; %s.\n", strerror(errno));
			exit(1);
		}
		size_t n = 0;
		mh_table_serialize(table, nullptr, 0, &n);
		std::vector<uint8_t> tb(n ? n : 1);
		if(mh_table_serialize(table, tb.data(), tb.size(), &n) != MH_OK || (n && fwrite(tb.data(), 1, n, ef) != n)) {
			eprintf("Error occurred while writing file.\n");
			exit(1);
		}
		fclose(ef);
	}

	if(extract) {
		eprintf("Extracting %s ===> %s...\n", input, shown(output));
		uint64_t n_out = 0;
		const int gpus = gpus_for(in_bytes.n, 5);
		bool done = false;
		if(gpus > 1) {   // the payload is cut at arbitrary bits: speculative start per GPU, seam handshake
			if(!result.resize(in_bytes.n * 4 + 4096)) alloc_fail();
			for(int attempt = 0; attempt < 2 && !done; ++attempt) {
				std::vector<uint64_t> lens(size_t(gpus), 0);
				rc = on_every_rank(gpus, [&](int r, mh_comm* c) {
					return mh_sharded_decompress_host(c, table, in_bytes.p, in_bytes.n, result.p, result.cap, &lens[size_t(r)]);
				});
				n_out = lens[0];
				if(rc == MH_OK || rc == MH_ERR_CORRUPT_STREAM) done = true;
				else if(rc == MH_ERR_CAPACITY && n_out > result.cap) { if(!result.resize(n_out)) alloc_fail(); }   // now the size is known
				else if(rc == MH_ERR_BAD_HEADER || rc == MH_ERR_TYPE_MISMATCH) die_status("extracting", rc);
				else break;   // too short to cut, or no second device after all: one GPU
			}
		}
		if(!done) {
			open_session();
			rc = mh_session_decompress(session, table, in_bytes.p, in_bytes.n, nullptr, 0, &n_out);   // decode, learn the size
			if(rc != MH_OK && rc != MH_ERR_CORRUPT_STREAM) die_status("extracting", rc);
			if(!result.resize(n_out)) alloc_fail();
			rc = mh_session_fetch(session, result.p, result.cap, &n_out);
			if(rc == MH_ERR_WORKSPACE) {   // the stream did not fit the device buffers: the first pass only counted; decode again, chunk by chunk, into the host buffer
				rc = mh_session_decompress(session, table, in_bytes.p, in_bytes.n, result.p, result.cap, &n_out);
				if(rc != MH_OK && rc != MH_ERR_CORRUPT_STREAM) die_status("extracting", rc);
			} else if(rc != MH_OK) {
				die_status("extracting", rc);
			}
		}
		result.n = n_out;
		if(n_out && fwrite(result.p, 1, n_out, output_fd) != n_out) {
			eprintf("Error occurred while writing file.\n");
			exit(1);
		}
	} else {
		eprintf("Compressing %s ===> %s...\n", input, shown(output));
		if(!built_here) {
			open_session();
			if(!result.resize(in_bytes.n + in_bytes.n / 8 + 4200)) alloc_fail();
			uint64_t out_len = 0, dropped = 0;
			rc = mh_session_compress_with_table(session, table, in_bytes.p, in_bytes.n, result.p, result.n, &out_len, &dropped);
			if(rc == MH_ERR_CAPACITY && out_len > result.n) {   // a foreign table that expands this input: the call reported the size it needs
				if(!result.resize(out_len)) alloc_fail();
				rc = mh_session_compress_with_table(session, table, in_bytes.p, in_bytes.n, result.p, result.n, &out_len, &dropped);
			}
			if(rc != MH_OK) die_status("compressing", rc);
			result.n = out_len;
			// The reference drops a symbol that has no codeword in the given table without a word (its assert is compiled out,
			// src/coding.cpp:72, Makefile:20); the bytes written are the same, but say so.
			if(dropped)
				eprintf("Warning: %llu input byte%s no codeword in the provided encoding table and %s dropped; the output will not extract to the input.\n",
				        (unsigned long long) dropped, dropped == 1 ? " has" : "s have", dropped == 1 ? "was" : "were");
		}
		// The reference writes a 0x80 placeholder first and seeks back to patch the header (src/coding.cpp:66, :85-89).
		// On a non-seekable sink the seek fails and the header byte lands after the payload; keep that behaviour.
		const bool seekable = fseek(output_fd, 0, SEEK_CUR) == 0;
		bool ok;
		if(seekable) {
			ok = fwrite(result.p, 1, result.n, output_fd) == result.n;
		} else {
			const uint8_t placeholder = 0x80;
			ok = fwrite(&placeholder, 1, 1, output_fd) == 1 &&
			     (result.n == 1 || fwrite(result.p + 1, 1, result.n - 1, output_fd) == result.n - 1) &&
			     fwrite(result.p, 1, 1, output_fd) == 1;
		}
		if(!ok) {
			eprintf("Error occurred while writing file.\n");
			exit(1);
		}
	}
	if(output_fd != stdout) fclose(output_fd);
	else fflush(stdout);

	mh_table_destroy(table);
	mh_session_destroy(session);
	eprintf("Done.\n");
	return 0;
}
