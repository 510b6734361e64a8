// mh_synth.cu — device generators for the benchmark workloads. Not part of the reference.
// Definitions are shared with oracle/mh_oracle.c (mho_synth_markov / mho_synth_fibonacci) and the tests compare
// the two byte for byte:
//   rnd(i)  = splitmix64 finaliser of (seed + (i + 1) * 0x9E3779B97F4A7C15),  i = global byte index
//   target  = (hi32(rnd) * row_total) >> 32
//   symbol  = smallest c with cumulative_count[c] > target
#include <vector>

#include "mh_internal.hpp"

namespace mh {

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
	z ^= z >> 27; z *= 0x94D049BB133111EBull;
	z ^= z >> 31;
	return z;
}
__device__ __forceinline__ uint64_t rnd_at(uint64_t seed, uint64_t i) { return mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ull); }

// Compact rows: for context p the live successors are syms[row_begin[p] .. row_begin[p+1]) with inclusive
// cumulative counts cum[...]; total[p] = last cum of the row (0 if the row is empty).
__global__ void synth_markov_kernel(const uint32_t* __restrict__ row_begin, const uint8_t* __restrict__ syms,
                                    const uint64_t* __restrict__ cum, const uint64_t* __restrict__ total, uint64_t seed,
                                    uint64_t seg_bytes, uint64_t first_seg, uint8_t* __restrict__ out, uint64_t n) {
	const uint64_t s = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
	const uint64_t off = s * seg_bytes;
	if(off >= n) return;
	const uint64_t len = n - off < seg_bytes ? n - off : seg_bytes;
	const uint64_t g0 = (first_seg + s) * seg_bytes;
	uint32_t prev = ' ';
	for(uint64_t j = 0; j < len; ++j) {
		const uint32_t row = total[prev] ? prev : uint32_t(' ');
		uint32_t sym = ' ';
		const uint64_t tot = total[row];
		if(tot) {
			const uint64_t target = ((rnd_at(seed, g0 + j) >> 32) * tot) >> 32;
			uint32_t lo = row_begin[row], hi = row_begin[row + 1] - 1;   // first index with cum > target
			while(lo < hi) {
				const uint32_t mid = (lo + hi) >> 1;
				if(cum[mid] > target) hi = mid; else lo = mid + 1;
			}
			sym = syms[lo];
		}
		out[off + j] = uint8_t(sym);
		prev = sym;
	}
}

__global__ void synth_fib_kernel(int k, uint32_t base, uint64_t seed, uint64_t first_index, uint8_t* __restrict__ out, uint64_t n) {
	__shared__ uint64_t cum[64];
	if(threadIdx.x == 0) {
		uint64_t a = 1, b = 1, s = 0;
		for(int j = 0; j < k; ++j) { s += a; cum[j] = s; const uint64_t nx = a + b; a = b; b = nx; }
	}
	__syncthreads();
	const uint64_t tot = cum[k - 1];
	for(uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
		const uint64_t target = ((rnd_at(seed, first_index + i) >> 32) * tot) >> 32;
		int lo = 0, hi = k - 1;
		while(lo < hi) {
			const int mid = (lo + hi) >> 1;
			if(cum[mid] > target) hi = mid; else lo = mid + 1;
		}
		out[i] = uint8_t(base + lo);
	}
}

}  // namespace
}  // namespace mh

extern "C" int mh_synth_markov(const uint32_t* trans_counts, uint64_t seed, uint64_t seg_bytes, uint64_t first_seg,
                               uint8_t* d_out, uint64_t n, mh_stream_t stream) {
	using namespace mh;
	if(!trans_counts || !d_out || seg_bytes == 0) return MH_ERR_INVALID_ARG;
	if(n == 0) return MH_OK;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	std::vector<uint32_t> row_begin(257);
	std::vector<uint8_t> syms;
	std::vector<uint64_t> cum, total(256, 0);
	for(int p = 0; p < 256; ++p) {
		row_begin[p] = uint32_t(syms.size());
		uint64_t s = 0;
		for(int c = 0; c < 256; ++c)
			if(trans_counts[256 * p + c]) { s += trans_counts[256 * p + c]; syms.push_back(uint8_t(c)); cum.push_back(s); }
		total[p] = s;
	}
	row_begin[256] = uint32_t(syms.size());
	if(syms.empty()) { syms.push_back(' '); cum.push_back(0); }
	uint32_t* d_rows; uint8_t* d_syms; uint64_t* d_cum; uint64_t* d_total;
	MH_CUDA(cudaMalloc(&d_rows, 257 * 4));
	MH_CUDA(cudaMalloc(&d_syms, syms.size()));
	MH_CUDA(cudaMalloc(&d_cum, cum.size() * 8));
	MH_CUDA(cudaMalloc(&d_total, 256 * 8));
	MH_CUDA(cudaMemcpyAsync(d_rows, row_begin.data(), 257 * 4, cudaMemcpyHostToDevice, st));
	MH_CUDA(cudaMemcpyAsync(d_syms, syms.data(), syms.size(), cudaMemcpyHostToDevice, st));
	MH_CUDA(cudaMemcpyAsync(d_cum, cum.data(), cum.size() * 8, cudaMemcpyHostToDevice, st));
	MH_CUDA(cudaMemcpyAsync(d_total, total.data(), 256 * 8, cudaMemcpyHostToDevice, st));
	const uint64_t segs = (n + seg_bytes - 1) / seg_bytes;
	synth_markov_kernel<<<unsigned((segs + 127) / 128), 128, 0, st>>>(d_rows, d_syms, d_cum, d_total, seed, seg_bytes, first_seg, d_out, n);
	count_launch(1);
	MH_CUDA(cudaGetLastError());
	MH_CUDA(cudaStreamSynchronize(st));   // the host vectors above must outlive the copies
	cudaFree(d_rows); cudaFree(d_syms); cudaFree(d_cum); cudaFree(d_total);
	return MH_OK;
}

extern "C" int mh_synth_fibonacci(int k_symbols, uint8_t base, uint64_t seed, uint64_t first_index, uint8_t* d_out,
                                  uint64_t n, mh_stream_t stream) {
	using namespace mh;
	if(!d_out || k_symbols < 1 || k_symbols > 64) return MH_ERR_INVALID_ARG;
	if(n == 0) return MH_OK;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint64_t blocks = (n + 255) / 256;
	synth_fib_kernel<<<unsigned(blocks < 65535 * 16 ? blocks : 65535 * 16), 256, 0, st>>>(k_symbols, base, seed, first_index, d_out, n);
	count_launch(1);
	MH_CUDA(cudaGetLastError());
	return MH_OK;
}
