// mh_histogram.cu — kernel 1: order-0 / order-1 context histogram for sm_100a.
//
// Replaces construct_table and its two counting lambdas (reference src/main.cpp:29-39, :168-170, :176-178):
//   order 1: counts[256*prev + c]++ with prev seeded by prev0;  order 0: counts[c]++.
//
// Shape of the kernel:
//   * every thread streams 16 input bytes per 128-bit coalesced load (ld.global.nc, no L1 allocation); the byte
//     before a thread's 16 comes from the neighbouring lane by shuffle, so each input byte is read once;
//   * bins live in privatised shared-memory sub-histograms. The full 256x256 u32 table (256 KiB) does not fit
//     one SM, so a probe kernel first finds the byte range [lo, lo+R) that the data actually uses (text: R ~ 113)
//     and the main kernel keeps an RxR box in shared memory, replicated per warp group when it is small;
//     pairs outside the box (rare for text, everything beyond R for binary data) go straight to the global
//     64-bit table with L2 atomics, so the result is exact whatever the probe saw;
//   * the boxes are folded into the global table with one 64-bit atomic per non-zero bin.
#include "mh_internal.hpp"

namespace mh {

namespace {

constexpr int kHistThreads = 512;
constexpr int kHistWarps = kHistThreads / 32;
constexpr int kProbeThreads = 256;
constexpr int kProbeWindows = 64;
constexpr int kProbeWindowBytes = 4096;

__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}

// Probe: kProbeWindows CTAs each mark the byte values present in one 4 KiB window spread over the input; the
// result is a 256-bit presence bitmap in params[0..7] (zeroed by the launcher).
__global__ void __launch_bounds__(kProbeThreads) hist_probe_kernel(const uint8_t* __restrict__ in, uint64_t n,
                                                                    uint32_t* __restrict__ params) {
	const uint64_t windows = n < uint64_t(kProbeWindows) * kProbeWindowBytes ? 1 : kProbeWindows;
	if(blockIdx.x >= windows) return;
	const uint64_t stride = windows > 1 ? (n - kProbeWindowBytes) / (windows - 1) : 0;
	const uint64_t wbytes = windows > 1 ? kProbeWindowBytes : n;
	uint32_t mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	for(uint64_t i = threadIdx.x; i < wbytes; i += kProbeThreads) {
		const uint32_t b = in[blockIdx.x * stride + i];
#pragma unroll
		for(int k = 0; k < 8; ++k) mine[k] |= (b >> 5) == uint32_t(k) ? (1u << (b & 31)) : 0u;
	}
#pragma unroll
	for(int k = 0; k < 8; ++k) {
		const uint32_t v = __reduce_or_sync(0xffffffffu, mine[k]);
		if((threadIdx.x & 31) == 0 && v) atomicOr(&params[k], v);
	}
}

// From the presence bitmap: the byte range [lo, lo + R) to keep in shared memory and how often to replicate it.
__device__ __forceinline__ void hist_plan(const uint32_t* __restrict__ params, uint32_t smem_words, uint32_t& lo_out,
                                          uint32_t& r_out, uint32_t& reps_out) {
	int lo = 256, hi = -1;
	for(int k = 0; k < 8; ++k) {
		const uint32_t w = params[k];
		if(!w) continue;
		if(lo == 256) lo = 32 * k + (__ffs(w) - 1);
		hi = 32 * k + (31 - __clz(w));
	}
	if(hi < 0) { lo = 0; hi = 0; }
	uint32_t range = uint32_t(hi - lo + 1);
	uint32_t rmax = 1;
	while((rmax + 1) * (rmax + 1) <= smem_words) ++rmax;
	if(range > rmax) range = rmax;
	uint32_t reps = smem_words / (range * range);
	if(reps > uint32_t(kHistWarps)) reps = kHistWarps;
	lo_out = uint32_t(lo);
	r_out = range;
	reps_out = reps;
}

template <int ORDER>
__global__ void __launch_bounds__(kHistThreads) hist_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t prev0,
                                                             unsigned long long* __restrict__ counts,
                                                             const uint32_t* __restrict__ params, uint32_t smem_words) {
	extern __shared__ uint32_t sh[];
	__shared__ uint32_t s_plan[3];
	if(ORDER) {
		if(threadIdx.x == 0) hist_plan(params, smem_words, s_plan[0], s_plan[1], s_plan[2]);
		__syncthreads();
	}
	const uint32_t lo = ORDER ? s_plan[0] : 0u;
	const uint32_t R = ORDER ? s_plan[1] : 256u;
	const uint32_t reps = ORDER ? s_plan[2] : uint32_t(kHistWarps);
	const uint32_t box = ORDER ? R * R : 256u;
	for(uint32_t i = threadIdx.x; i < reps * box; i += kHistThreads) sh[i] = 0;
	__syncthreads();
	uint32_t* mine = sh + ((threadIdx.x >> 5) % reps) * box;
	const uint32_t lane = threadIdx.x & 31;

	auto tally = [&](uint32_t p, uint32_t c) {
		if(ORDER) {
			const uint32_t up = p - lo, uc = c - lo;
			if(up < R && uc < R) atomicAdd(&mine[up * R + uc], 1u);
			else atomicAdd(&counts[p * 256u + c], 1ull);
		} else {
			atomicAdd(&mine[c], 1u);
		}
	};

	// [0, head) unaligned lead-in, then G aligned 16-byte groups, then the tail
	const uint64_t addr = reinterpret_cast<uint64_t>(in);
	uint64_t head = (16 - (addr & 15)) & 15;
	if(head > n) head = n;
	const uint64_t groups = (n - head) >> 4;
	const uint64_t gstride = uint64_t(gridDim.x) * kHistThreads;
	// warp-uniform trip count so the shuffle below always has all 32 lanes
	for(uint64_t base = uint64_t(blockIdx.x) * kHistThreads + (threadIdx.x & ~31u); base < groups; base += gstride) {
		const uint64_t g = base + lane;
		const bool live = g < groups;
		uint4 v = make_uint4(0, 0, 0, 0);
		if(live) v = ld_stream_128(in + head + (g << 4));
		uint32_t prev = __shfl_up_sync(0xffffffffu, v.w >> 24, 1);
		if(lane == 0) prev = (head + (g << 4)) == 0 ? prev0 : uint32_t(in[head + (g << 4) - 1]);
		if(live) {
			const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
			for(int k = 0; k < 4; ++k) {
#pragma unroll
				for(int b = 0; b < 4; ++b) {
					const uint32_t c = (w[k] >> (8 * b)) & 255u;
					tally(prev, c);
					prev = c;
				}
			}
		}
	}
	if(blockIdx.x == 0 && threadIdx.x == 0) {
		uint32_t prev = prev0;
		for(uint64_t i = 0; i < head; ++i) { tally(prev, in[i]); prev = in[i]; }
		const uint64_t t0 = head + (groups << 4);
		if(t0 < n) {
			prev = t0 == 0 ? prev0 : uint32_t(in[t0 - 1]);
			for(uint64_t i = t0; i < n; ++i) { tally(prev, in[i]); prev = in[i]; }
		}
	}
	__syncthreads();
	for(uint32_t i = threadIdx.x; i < reps * box; i += kHistThreads) {
		const uint32_t c = sh[i];
		if(!c) continue;
		const uint32_t bin = i % box;
		if(ORDER) atomicAdd(&counts[(bin / R + lo) * 256u + (bin % R + lo)], (unsigned long long) c);
		else atomicAdd(&counts[bin], (unsigned long long) c);
	}
}

}  // namespace

int launch_histogram(const uint8_t* d_in, uint64_t n, uint8_t prev0, int order, unsigned long long* d_counts,
                     mh_workspace* ws, cudaStream_t st) {
	if(order != 0 && order != 1) return MH_ERR_INVALID_ARG;
	if(!d_counts || (!d_in && n)) return MH_ERR_INVALID_ARG;
	if(!ws || !ws->hist_params) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * (order ? 65536 : 256), st));
	if(n == 0) return MH_OK;
	// each CTA counts in u32: keep a CTA's share below 2^32 samples
	const int sms = sm_count();
	const uint64_t groups = n / 16 + 1;
	uint64_t want = (groups + kHistThreads - 1) / kHistThreads;
	int ctas_per_sm = 2;
	uint64_t grid = uint64_t(sms) * ctas_per_sm;
	if(grid > want) grid = want;
	while(n / grid >= (1ull << 32)) grid *= 2;
	if(order) {
		const uint32_t smem_bytes = 100 * 1024;
		static bool attr_done = false;
		if(!attr_done) {
			MH_CUDA(cudaFuncSetAttribute(hist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
			attr_done = true;
		}
		{
			MH_CUDA(cudaMemsetAsync(ws->hist_params, 0, 8 * sizeof(uint32_t), st));
			ProfScope p("hist_probe_kernel", st);
			hist_probe_kernel<<<kProbeWindows, kProbeThreads, 0, st>>>(d_in, n, ws->hist_params);
		}
		{
			ProfScope p("hist_kernel<1>", st);
			hist_kernel<1><<<unsigned(grid), kHistThreads, smem_bytes, st>>>(d_in, n, prev0, d_counts, ws->hist_params, smem_bytes / 4);
		}
		count_launch(2);
	} else {
		ProfScope p("hist_kernel<0>", st);
		hist_kernel<0><<<unsigned(grid), kHistThreads, kHistWarps * 256 * 4, st>>>(d_in, n, prev0, d_counts, ws->hist_params, 0u);
		count_launch(1);
	}
	MH_CUDA(cudaGetLastError());
	return MH_OK;
}

}  // namespace mh
