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
//
// Lane-private variant (hist_lane_kernel, the one text-like inputs take). 32 lanes that add into 32 random bins hit
// the same shared-memory bank ~3.3 deep, and that serialisation — not HBM — bounded the box kernel (ncu:
// 3.3 wavefronts per ATOMS). When the probe finds an alphabet of K <= 53 byte values, every lane of a warp gets
// its OWN column of bins: dense pair index (rank(prev), rank(c)) x 32 lanes, u16 counters packed in pairs so that
// the word of (bin, lane) sits in bank `lane` — one wavefront per warp-wide atomic, whatever the data. The rank
// comes from a lane-replicated u32 lookup (also conflict-free) that already carries the row offset, the column
// offset and the half-word selector. u16 counters cannot overflow because the CTA drains them into the global
// table every kLaneFlushIters iterations. Alphabets up to 221 values use the same code with u32 bins in 16..1
// lane columns. Byte values the probe did not see land on a trash row/column and are fixed up per 16-byte group
// with global atomics, so the result stays exact whatever the probe saw.
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

// ---- lane-private kernels ---------------------------------------------------------------------------------
constexpr int kLaneThreads = 1024;
constexpr int kLaneWarps = kLaneThreads / 32;
constexpr uint32_t kLaneSmemBytes = 224 * 1024;       // dynamic shared memory of hist_lane_kernel
constexpr uint32_t kLaneLutBytes = 256 * 32 * 4;      // lane-replicated symbol lookup
constexpr uint32_t kLaneFlushIters = 63;              // 63 iterations x 32 warps x 2 groups x 16 bytes < 65536

__device__ __forceinline__ void red_add_shared(uint32_t addr, uint32_t v) {
	asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

// mode 0: general box kernel; 1: packed u16 x 32 lane columns; 2: u32 x S lane columns (S = 16, 8, 4, 2, 1)
struct LanePlan {
	uint32_t mode, K, pitch, S;
};
__device__ __forceinline__ LanePlan lane_plan(const uint32_t* __restrict__ params) {
	LanePlan p;
	uint32_t k = 0;
	for(int i = 0; i < 8; ++i) k += __popc(params[i]);
	p.K = k;
	const uint32_t k1 = k + 1, room = kLaneSmemBytes - kLaneLutBytes - 128;   // bin rows are 128-byte aligned
	p.pitch = (k1 + 1) & ~1u;
	p.S = 32;
	p.mode = 0;
	if(k1 * p.pitch * 64u <= room) { p.mode = 1; return p; }
	for(uint32_t s = 16; s >= 1; s >>= 1)
		if(k1 * k1 * 4u * s <= room) { p.mode = 2; p.S = s; p.pitch = k1; return p; }
	return p;
}

// Lookup word of byte value c in lane l's column (rank ic, absolute shared-space address of its bin row):
//   [31:20] byte offset of (column ic, lane l) inside a bin row
//   [19:7] / [19:2]  shared-space address of context ic's bin row (128- / 4-byte aligned)
//   packed mode: [4:0] shift of the u16 half (0 or 16), [5] trash;   u32 mode: [1] trash
template <bool PACK16>
struct LaneFmt {
	static constexpr uint32_t kRowMask = PACK16 ? 0xfff80u : 0xffffcu;
	static constexpr uint32_t kTrash = PACK16 ? 0x20u : 0x2u;
};

// 16 consecutive bytes whose predecessor is `prev`: one conflict-free lookup and one conflict-free atomic per byte.
template <bool PACK16>
__device__ __forceinline__ uint32_t lane_tally16(const uint4 v, uint32_t prev, uint32_t lutl) {
	uint32_t e = lds_u32(lutl + prev * 128u);
	uint32_t acc = e;
	uint32_t row = e & LaneFmt<PACK16>::kRowMask;
	const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
	for(int k = 0; k < 4; ++k) {
#pragma unroll
		for(int b = 0; b < 4; ++b) {
			const uint32_t c = __byte_perm(w[k], 0, 0x4440 + b);
			e = lds_u32(lutl + c * 128u);
			red_add_shared(row + (e >> 20), PACK16 ? __funnelshift_l(0u, 1u, e) : 1u);
			row = e & LaneFmt<PACK16>::kRowMask;
			acc |= e;
		}
	}
	return acc;
}

template <int THREADS>
__device__ __forceinline__ void hist_box_body(const uint8_t* __restrict__ in, uint64_t n, uint32_t prev0,
                                              unsigned long long* __restrict__ counts, const uint32_t* __restrict__ params,
                                              uint32_t smem_words, uint32_t* sh);

__global__ void __launch_bounds__(kLaneThreads, 1) hist_lane_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t prev0,
                                                                     unsigned long long* __restrict__ counts,
                                                                     const uint32_t* __restrict__ params) {
	extern __shared__ uint32_t sh[];   // [256 x 32] lookup | bins
	__shared__ uint8_t s_idx[256];     // byte value -> rank (K for values the probe did not see)
	__shared__ uint8_t s_inv[256];     // rank -> byte value
	__shared__ LanePlan s_plan;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if(tid == 0) s_plan = lane_plan(params);
	if(tid < 256) {
		const uint32_t word = tid >> 5, bit = tid & 31;
		uint32_t rank = 0;
		for(uint32_t i = 0; i < word; ++i) rank += __popc(params[i]);
		rank += __popc(params[word] & ((1u << bit) - 1u));
		const bool present = (params[word] >> bit) & 1u;
		s_idx[tid] = 0xff;
		if(present) { s_idx[tid] = uint8_t(rank); s_inv[rank] = uint8_t(tid); }
	}
	__syncthreads();
	const LanePlan P = s_plan;
	if(P.mode == 0) {   // alphabet too large for lane-private bins: the R x R box path, in this same launch
		hist_box_body<kLaneThreads>(in, n, prev0, counts, params, kLaneSmemBytes / 4, sh);
		return;
	}
	const bool pack16 = P.mode == 1;
	const uint32_t K = P.K, K1 = K + 1;
	const uint32_t row_bytes = pack16 ? (P.pitch / 2) * 128u : K1 * P.S * 4u;
	const uint32_t bin_words = K1 * row_bytes / 4;
	if(tid < 256 && s_idx[tid] == 0xff) s_idx[tid] = uint8_t(K);
	__syncthreads();
	// bin rows start on a 128-byte boundary: the lookup word keeps flags in the low address bits
	const uint32_t bins_sa = (uint32_t(__cvta_generic_to_shared(sh)) + kLaneLutBytes + 127u) & ~127u;
	uint32_t* bins = sh + (bins_sa - uint32_t(__cvta_generic_to_shared(sh))) / 4;
	for(uint32_t i = tid; i < 256 * 32; i += kLaneThreads) {
		const uint32_t c = i >> 5, l = i & 31, ic = s_idx[c];
		const uint32_t col = pack16 ? (ic >> 1) * 128u + l * 4u : (ic * P.S + (l & (P.S - 1))) * 4u;
		const uint32_t flags = pack16 ? ((ic & 1u) << 4) | (ic == K ? LaneFmt<true>::kTrash : 0u) : (ic == K ? LaneFmt<false>::kTrash : 0u);
		sh[i] = (col << 20) | (bins_sa + ic * row_bytes) | flags;
	}
	for(uint32_t i = tid; i < bin_words; i += kLaneThreads) bins[i] = 0;
	__syncthreads();
	const uint32_t lutl = uint32_t(__cvta_generic_to_shared(sh)) + lane * 4;

	// packed mode: drain the u16 counters into the global table (warp per 32-lane word row)
	auto drain16 = [&]() {
		__syncthreads();
		const uint32_t half = P.pitch / 2;
		for(uint32_t r = warp; r < K * half; r += kLaneWarps) {   // rows of the trash context are dropped
			const uint32_t v = bins[r * 32 + lane];
			bins[r * 32 + lane] = 0;
			const uint32_t lo = __reduce_add_sync(0xffffffffu, v & 0xffffu), hi = __reduce_add_sync(0xffffffffu, v >> 16);
			const uint32_t ip = r / half, ic = 2 * (r - ip * half);
			if(lane == 0 && lo && ic < K) atomicAdd(&counts[uint32_t(s_inv[ip]) * 256u + s_inv[ic]], (unsigned long long) lo);
			if(lane == 1 && hi && ic + 1 < K) atomicAdd(&counts[uint32_t(s_inv[ip]) * 256u + s_inv[ic + 1]], (unsigned long long) hi);
		}
		for(uint32_t i = K * half * 32 + tid; i < bin_words; i += kLaneThreads) bins[i] = 0;
		__syncthreads();
	};

	// [0, head) lead-in (1..16 bytes, so every 16-byte group has a predecessor byte in memory), G aligned groups, tail.
	// Each CTA takes one contiguous span of groups and walks it two CTA-wide rows (2 x 1024 groups) per iteration.
	const uint64_t addr = reinterpret_cast<uint64_t>(in);
	uint64_t head = 16 - (addr & 15);
	if(head > n) head = n;
	const uint64_t groups = (n - head) >> 4;
	const uint8_t* body = in + head;
	constexpr uint32_t kRow = kLaneThreads, kIterGroups = 2 * kRow;
	uint64_t span = (groups + gridDim.x - 1) / gridDim.x;
	span = (span + kIterGroups - 1) / kIterGroups * kIterGroups;
	const uint64_t g0 = uint64_t(blockIdx.x) * span < groups ? uint64_t(blockIdx.x) * span : groups;
	const uint64_t g1 = g0 + span < groups ? g0 + span : groups;
	const uint32_t full = uint32_t((g1 - g0) / kIterGroups);
	const uint32_t rest = uint32_t((g1 - g0) - uint64_t(full) * kIterGroups);

	// exact fix-up of a 16-byte group that touched the trash row/column (a byte value the probe missed)
	auto fix_up = [&](const uint4 v, uint32_t prev) {
		const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for(int k = 0; k < 4; ++k) {
#pragma unroll 1
			for(int b = 0; b < 4; ++b) {
				const uint32_t c = (w[k] >> (8 * b)) & 255u;
				if(s_idx[prev] == K || s_idx[c] == K) atomicAdd(&counts[prev * 256u + c], 1ull);
				prev = c;
			}
		}
	};
	auto process = [&](const uint4 v, uint32_t first_prev, bool live) {
		uint32_t prev = __shfl_up_sync(0xffffffffu, v.w >> 24, 1);
		if(lane == 0) prev = first_prev;
		if(live) {
			const uint32_t acc = pack16 ? lane_tally16<true>(v, prev, lutl) & LaneFmt<true>::kTrash : lane_tally16<false>(v, prev, lutl) & LaneFmt<false>::kTrash;
			if(acc) fix_up(v, prev);
		}
	};
	auto load2 = [&](const uint8_t* q, uint4& x, uint4& y, uint32_t& px, uint32_t& py) {
		x = ld_stream_128(q);
		y = ld_stream_128(q + kRow * 16);
		if(lane == 0) { px = q[-1]; py = q[kRow * 16 - 1]; }
	};

	const uint8_t* p = body + ((g0 + tid) << 4);
	uint4 a = make_uint4(0, 0, 0, 0), b = a;
	uint32_t pa = 0, pb = 0, iters = 0;
	if(full) load2(p, a, b, pa, pb);
	// ptxas gives all these loads ONE scoreboard: issued at the top of the trip, the loads of the next two rows made the first
	// use of the current rows wait for THEM — a full memory latency per trip, no prefetch at all (28 % of the kernel's stall
	// samples sat on that one instruction, profiles/r02_hist_lane_kernel.txt). So the next rows' address depends on the
	// current rows' registers (through a zero ptxas cannot see): the wait comes first, then the loads, then ~260 instructions
	// of counting in which nothing waits for memory.
	const uint32_t opaque_zero = uint32_t(n >> 63);
	for(uint32_t i = 0; i < full; ++i) {   // the next two rows are in flight while the current two are counted
		uint4 na = make_uint4(0, 0, 0, 0), nb = na;
		uint32_t npa = 0, npb = 0;
		const uint32_t gate = (a.w ^ b.w ^ pa ^ pb) & opaque_zero;
		if(i + 1 < full) load2(p + kIterGroups * 16 + gate, na, nb, npa, npb);
		process(a, pa, true);
		process(b, pb, true);
		a = na; b = nb; pa = npa; pb = npb;
		p += kIterGroups * 16;
		if(pack16 && ++iters == kLaneFlushIters) { iters = 0; drain16(); }   // CTA-uniform trip count
	}
	if(rest) {   // the span's ragged end (CTA-uniform condition; the shuffles need whole warps)
		const bool la = tid < rest, lb = tid + kRow < rest;
		a = b = make_uint4(0, 0, 0, 0);
		if(la) { a = ld_stream_128(p); if(lane == 0) pa = p[-1]; }
		if(lb) { b = ld_stream_128(p + kRow * 16); if(lane == 0) pb = p[kRow * 16 - 1]; }
		process(a, pa, la);
		process(b, pb, lb);
	}
	if(blockIdx.x == 0 && tid == 0) {   // unaligned lead-in and the ragged tail: a few bytes, straight to the global table
		uint32_t prev = prev0;
		for(uint64_t i = 0; i < head; ++i) { atomicAdd(&counts[prev * 256u + in[i]], 1ull); prev = in[i]; }
		const uint64_t t0 = head + (groups << 4);
		if(t0 < n) {
			prev = t0 == 0 ? prev0 : uint32_t(in[t0 - 1]);
			for(uint64_t i = t0; i < n; ++i) { atomicAdd(&counts[prev * 256u + in[i]], 1ull); prev = in[i]; }
		}
	}
	if(pack16) {
		drain16();
	} else {
		__syncthreads();
		for(uint32_t bin = tid; bin < K * K1; bin += kLaneThreads) {
			const uint32_t ip = bin / K1, ic = bin - ip * K1;
			if(ic == K) continue;
			uint32_t s = 0;
			for(uint32_t j = 0; j < P.S; ++j) s += bins[bin * P.S + j];
			if(s) atomicAdd(&counts[uint32_t(s_inv[ip]) * 256u + s_inv[ic]], (unsigned long long) s);
		}
	}
}

// order 0: 256 bins x 32 lane columns (32 KiB): bank = lane, no conflicts
__global__ void __launch_bounds__(kLaneThreads, 2) hist0_lane_kernel(const uint8_t* __restrict__ in, uint64_t n,
                                                                      unsigned long long* __restrict__ counts) {
	__shared__ uint32_t bins[256 * 32];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for(uint32_t i = tid; i < 256 * 32; i += kLaneThreads) bins[i] = 0;
	__syncthreads();
	const uint32_t mine = uint32_t(__cvta_generic_to_shared(bins)) + lane * 4;
	const uint64_t addr = reinterpret_cast<uint64_t>(in);
	uint64_t head = (16 - (addr & 15)) & 15;
	if(head > n) head = n;
	const uint64_t groups = (n - head) >> 4;
	const uint8_t* body = in + head;
	const uint64_t gstride = uint64_t(gridDim.x) * kLaneThreads;
	auto tally16 = [&](const uint4 v) {
		const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for(int k = 0; k < 4; ++k)
#pragma unroll
			for(int b = 0; b < 4; ++b) red_add_shared(mine + (((w[k] >> (8 * b)) & 255u) << 7), 1u);
	};
	uint64_t g = uint64_t(blockIdx.x) * kLaneThreads + tid;
	uint4 a = make_uint4(0, 0, 0, 0), b = a;
	if(g < groups) a = ld_stream_128(body + (g << 4));
	if(g + gstride < groups) b = ld_stream_128(body + ((g + gstride) << 4));
	for(; g < groups; g += 2 * gstride) {   // (ordering the next rows' loads behind the first use of the current rows, as hist_lane_kernel does, measured 4 % slower here: two CTAs per SM hide the wait)
		uint4 na = make_uint4(0, 0, 0, 0), nb = na;
		if(g + 2 * gstride < groups) na = ld_stream_128(body + ((g + 2 * gstride) << 4));
		if(g + 3 * gstride < groups) nb = ld_stream_128(body + ((g + 3 * gstride) << 4));
		tally16(a);
		if(g + gstride < groups) tally16(b);
		a = na; b = nb;
	}
	if(blockIdx.x == 0 && tid == 0) {
		for(uint64_t i = 0; i < head; ++i) atomicAdd(&counts[in[i]], 1ull);
		for(uint64_t i = head + (groups << 4); i < n; ++i) atomicAdd(&counts[in[i]], 1ull);
	}
	__syncthreads();
	for(uint32_t r = warp; r < 256; r += kLaneWarps) {
		const uint32_t s = __reduce_add_sync(0xffffffffu, bins[r * 32 + lane]);
		if(lane == 0 && s) atomicAdd(&counts[r], (unsigned long long) s);
	}
}

// The general box path, called by hist_lane_kernel when the alphabet is too large for lane-private bins (all
// threads of the CTA enter; `sh` is the CTA's dynamic shared memory of smem_words words).
template <int THREADS>
__device__ __forceinline__ void hist_box_body(const uint8_t* __restrict__ in, uint64_t n, uint32_t prev0,
                                              unsigned long long* __restrict__ counts, const uint32_t* __restrict__ params,
                                              uint32_t smem_words, uint32_t* sh) {
	constexpr int ORDER = 1;
	constexpr int kHistThreads = THREADS;
	__shared__ uint32_t s_plan[3];
	if(threadIdx.x == 0) hist_plan(params, smem_words, s_plan[0], s_plan[1], s_plan[2]);
	__syncthreads();
	const uint32_t lo = s_plan[0];
	const uint32_t R = s_plan[1];
	const uint32_t reps = s_plan[2];
	const uint32_t box = R * R;
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

static std::atomic<uint64_t> g_hist_attr_done{0};

int launch_histogram(const uint8_t* d_in, uint64_t n, uint8_t prev0, int order, unsigned long long* d_counts,
                     mh_workspace* ws, cudaStream_t st, bool accumulate) {
	if(order != 0 && order != 1) return MH_ERR_INVALID_ARG;
	if(!d_counts || (!d_in && n)) return MH_ERR_INVALID_ARG;
	if(!ws || !ws->hist_params) return MH_ERR_WORKSPACE;
	if(!accumulate) MH_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * (order ? 65536 : 256), st));
	if(n == 0) return MH_OK;
	// each CTA counts in u32: keep a CTA's share below 2^32 samples
	const int sms = sm_count();
	const uint64_t groups = n / 16 + 1;
	auto grid_for = [&](int threads, int ctas_per_sm) {
		const uint64_t want = (groups + threads - 1) / threads;
		uint64_t grid = uint64_t(sms) * ctas_per_sm;
		if(grid > want) grid = want;
		while(n / grid >= (1ull << 32)) grid *= 2;
		return unsigned(grid);
	};
	if(order) {
		if(first_use_on_device(g_hist_attr_done))
			MH_CUDA(cudaFuncSetAttribute(hist_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kLaneSmemBytes)));
		{
			MH_CUDA(cudaMemsetAsync(ws->hist_params, 0, 8 * sizeof(uint32_t), st));
			ProfScope p("hist_probe_kernel", st);
			hist_probe_kernel<<<kProbeWindows, kProbeThreads, 0, st>>>(d_in, n, ws->hist_params);
		}
		// One counting kernel: it derives its plan (lane-private bins, or the R x R box for large alphabets) from the
		// probe's bitmap on the device, so no host round trip sits between the probe and the count.
		{
			ProfScope p("hist_lane_kernel", st);
			hist_lane_kernel<<<grid_for(kLaneThreads, 1), kLaneThreads, kLaneSmemBytes, st>>>(d_in, n, prev0, d_counts, ws->hist_params);
		}
		count_launch(2);
	} else {
		ProfScope p("hist0_lane_kernel", st);
		hist0_lane_kernel<<<grid_for(kLaneThreads, 2), kLaneThreads, 0, st>>>(d_in, n, d_counts);
		count_launch(1);
	}
	MH_CUDA(cudaGetLastError());
	return MH_OK;
}

}  // namespace mh
