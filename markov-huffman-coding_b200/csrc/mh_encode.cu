// mh_encode.cu — kernel 2: table-lookup encoder with a single-pass decoupled look-back scan and bit packing.
//
// Replaces the loop of i_coding_provider::compress (reference src/coding.cpp:67-78) together with the bit packer
// bitbuffer::push_encoding_descriptor / push_byte / flush (src/bitbuffer.cpp:21-73, :170-180): for every input
// byte c with predecessor prev, append codeword[prev][c] MSB-first to one contiguous bit stream.
//
// Shape of the kernel. Persistent CTAs (a multiple of the SM count) pull tiles of ROUNDS x 4 KiB of
// input from an atomic ticket, so a tile only ever waits on tiles that are already running.
//   0. once per CTA the codebook is staged in shared memory: u32 entries (27-bit left-aligned code | 5-bit length)
//      for the RxR box of byte values the table actually uses (text: R ~ 113 -> 50 KiB). Tables whose box does not
//      fit, or whose codewords exceed 27 bits, are gathered from global memory (L1/L2) instead;
//   1. per round, each thread takes 16 consecutive bytes with one coalesced 128-bit load (the byte before them
//      comes from the neighbouring lane by shuffle), looks up its 16 entries and sums their lengths;
//   2. a block-wide exclusive scan of the per-thread bit counts gives every thread its bit offset in the tile;
//   3. each thread funnels its codewords through a 64-bit window and ORs whole 32-bit words into a zeroed
//      shared-memory staging area (neighbouring threads share their boundary words, hence the OR);
//   4. warp 0 publishes the tile's (bit count, last 31 bits) in ONE 64-bit word and resolves the tile's global bit
//      offset by a warp-wide decoupled look-back (32 predecessors per poll, relaxed loads that bypass L1). The tile
//      that ends a partially filled 32-bit output word writes it, using the predecessor's published tail bits — so
//      every output word has exactly one writer: no pre-zeroed output, no global atomics, no second pass;
//   5. the staged bits are funnel-shifted by the tile's global bit phase and written with coalesced 32-bit
//      stores, byte-swapped so that stream bit p lands in byte p/8, bit 7 - p%8 (src/bitbuffer.cpp:12); the words
//      just copied are re-zeroed for the next tile.
#include "mh_internal.hpp"

namespace mh {

namespace {

constexpr int FMT_BOX_SMEM = 0;     // u32 box table in shared memory
constexpr int FMT_BOX_GLOBAL = 1;   // u32 box table gathered from global memory
constexpr int FMT_WIDE = 2;         // u64 entries (8-bit length | 56-bit code) gathered from global memory

constexpr uint64_t kAgg = 1ull << 62;   // aggregate word : kAgg | tail31 << 31 | bits31   (this tile only)
constexpr uint64_t kInc = 2ull << 62;   // inclusive word : kInc | bits62                   (all tiles up to this one)
constexpr uint64_t kLow31 = 0x7fffffffull;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
	for(int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
	return v;
}

// bits [pos, pos + count) of the MSB-first word array, count <= 32, right-aligned
__device__ __forceinline__ uint32_t stage_bits(const uint32_t* stage, uint32_t pos, uint32_t count) {
	if(count == 0) return 0;
	const uint32_t w = pos >> 5, off = pos & 31;
	const uint64_t two = (uint64_t(stage[w]) << 32) | stage[w + 1];
	return uint32_t((two << off) >> (64 - count));
}

// Warp-wide decoupled look-back over a window of kWin x 32 predecessors per poll. All 32 lanes of warp 0 call
// this. Returns, on every lane, the number of bits that precede this tile and the last min(that, 31) of them.
constexpr int kWin = 4;

__device__ __forceinline__ void look_back(uint32_t tile, uint32_t own_bits, uint32_t own_tail, unsigned long long* agg,
                                          unsigned long long* inc, unsigned long long& excl_bits, uint32_t& excl_tail) {
	const uint32_t lane = threadIdx.x & 31;
	if(lane == 0) st_relaxed(agg + tile, kAgg | (uint64_t(own_tail) << 31) | own_bits);
	unsigned long long sum = 0;
	unsigned long long nearest = kAgg;   // aggregate word of tile - 1 (an empty tile if there is none)
	bool first_window = true;
	long long base = (long long) tile - 1;
	while(base >= 0) {
		unsigned long long a[kWin], p[kWin];
#pragma unroll
		for(int k = 0; k < kWin; ++k) {
			const long long j = base - (k * 32 + int(lane));
			a[k] = kAgg; p[k] = kInc;   // tiles before the stream: inclusive, zero bits
			if(j >= 0) { p[k] = ld_relaxed(inc + j); a[k] = ld_relaxed(agg + j); }
		}
		int stop = -1;        // window index of the nearest predecessor with an inclusive prefix
		bool ready = true;    // every aggregate nearer than `stop` is visible
#pragma unroll
		for(int k = 0; k < kWin; ++k) {
			const unsigned has_inc = __ballot_sync(0xffffffffu, (p[k] >> 62) == 2);
			const unsigned has_agg = __ballot_sync(0xffffffffu, (a[k] >> 62) == 1);
			if(stop < 0) {
				unsigned need = 0xffffffffu;
				if(has_inc) { const int l = __ffs(has_inc) - 1; stop = k * 32 + l; need = (1u << l) - 1u; }
				if(k == 0 && first_window) need |= 1u;   // the tail always comes from tile - 1's aggregate word
				if((has_agg & need) != need) ready = false;
			}
		}
		if(!ready) continue;   // a predecessor has not published yet: poll again
		if(first_window) { nearest = __shfl_sync(0xffffffffu, a[0], 0); first_window = false; }
		unsigned long long v = 0;
#pragma unroll
		for(int k = 0; k < kWin; ++k) {
			const int idx = k * 32 + int(lane);
			if(stop < 0 || idx < stop) v += a[k] & kLow31;
			else if(idx == stop) v += p[k] & kDescValueMask;
		}
		sum += warp_sum(v);
		if(stop >= 0) break;
		base -= kWin * 32;
	}
	excl_bits = sum;
	uint32_t t_bits = uint32_t(nearest & kLow31), t_tail = uint32_t((nearest >> 31) & kLow31);
	if(t_bits < 31 && tile > 1) {
		// Rare (a predecessor produced fewer than 31 bits, e.g. dropped symbols): gather the tail serially.
		if(lane == 0) {
			for(long long j = (long long) tile - 2; j >= 0 && t_bits < 31; --j) {
				unsigned long long w;
				do { w = ld_relaxed(agg + j); } while((w >> 62) != 1);
				const uint32_t b = uint32_t(w & kLow31), tl = uint32_t((w >> 31) & kLow31);
				// concatenate (b, tl) in front of (t_bits, t_tail), keep the last 31 bits
				t_tail = uint32_t(((uint64_t(tl) << t_bits) | t_tail) & kLow31);
				t_bits = b + t_bits > 31 ? 31 : b + t_bits;
			}
		}
		t_tail = __shfl_sync(0xffffffffu, t_tail, 0);
	}
	excl_tail = t_tail;
	if(lane == 0) st_relaxed(inc + tile, kInc | (sum + own_bits));
}

struct EncArgs {
	const uint8_t* in;
	uint64_t n;
	uint32_t prev0;
	int order;
	const unsigned long long* wide;   // FMT_WIDE table
	const uint32_t* box;              // u32 box table with a zero border: (R + 1) x (R + 1), or 256 entries (order 0)
	uint32_t box_lo, box_r;
	uint32_t bit0;
	uint32_t stage_words;
	uint32_t* out_words;
	uint64_t out_capacity_words;
	unsigned long long* agg;
	unsigned long long* inc;
	uint32_t* ticket;
	unsigned long long* result;
	uint32_t n_tiles;
};

// One thread's 16 bytes -> 16 u32 entries (code left-aligned in [31:5], length in [4:0]). Returns the bit count;
// `floor` tracks the smallest entry seen (0 = some symbol had no codeword).
template <int FMT, bool FULL>
__device__ __forceinline__ uint32_t gather_box(const EncArgs& A, const uint32_t* table, const uint32_t (&w)[4], int live,
                                               uint32_t prev, uint32_t (&e)[16], uint32_t& floor) {
	const uint32_t R = A.box_r, lo = A.box_lo, pitch = R + 1;
	uint32_t bits = 0;
	uint32_t row = A.order ? min(prev - lo, R) * pitch : 0u;
#pragma unroll
	for(int i = 0; i < 16; ++i) {
		const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 255u;
		uint32_t ent = 0;
		if(FULL || i < live) {
			if(A.order) {
				const uint32_t uc = min(c - lo, R);   // bytes outside the box land on the zero border
				ent = FMT == FMT_BOX_SMEM ? table[row + uc] : __ldg(A.box + row + uc);
				row = uc * pitch;
			} else {
				ent = FMT == FMT_BOX_SMEM ? table[c] : __ldg(A.box + c);
			}
			floor = min(floor, ent);
		}
		e[i] = ent;
		bits += ent & 31u;
	}
	return bits;
}

template <int ROUNDS, int FMT, bool ALIGNED>
__global__ void __launch_bounds__(kEncThreads) encode_kernel(const EncArgs A) {
	extern __shared__ uint32_t smem[];
	uint32_t* stage = smem;                          // [stage_words + 4]
	uint32_t* table = smem + A.stage_words + 4;      // FMT_BOX_SMEM: [(R + 1)^2] or [256]
	__shared__ uint32_t warp_sums[kEncThreads / 32];
	__shared__ uint32_t s_tile[2];
	__shared__ unsigned long long s_prefix_bits;
	__shared__ uint32_t s_prefix_tail;

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for(uint32_t i = tid; i < A.stage_words + 4; i += kEncThreads) stage[i] = 0;
	if(FMT == FMT_BOX_SMEM) {
		const uint32_t entries = A.order ? (A.box_r + 1) * (A.box_r + 1) : 256u;
		for(uint32_t i = tid; i < entries; i += kEncThreads) table[i] = __ldg(A.box + i);
	}
	if(tid == 0) s_tile[0] = atomicAdd(A.ticket, 1u);
	uint32_t dropped = 0;

	for(uint32_t it = 0;; ++it) {
		__syncthreads();
		const uint32_t tile = s_tile[it & 1];
		if(tile >= A.n_tiles) break;
		if(tid == 0) s_tile[(it + 1) & 1] = atomicAdd(A.ticket, 1u);   // next ticket, off the critical path
		const uint64_t tile_base = uint64_t(tile) * (ROUNDS * kEncRoundBytes);
		uint32_t tile_bits = 0;   // running bit count of the tile (uniform across the block)
#pragma unroll 1
		for(int r = 0; r < ROUNDS; ++r) {
			const uint64_t round_base = tile_base + uint64_t(r) * kEncRoundBytes;
			if(round_base >= A.n) break;   // uniform
			const uint64_t my = round_base + tid * 16;
			// ---- 1. my 16 bytes + the byte before them ----
			uint32_t w[4] = {0, 0, 0, 0};
			int live = 0;
			if(my < A.n) {
				live = A.n - my >= 16 ? 16 : int(A.n - my);
				if(ALIGNED && live == 16) {
					const uint4 v = ld_stream_128(A.in + my);
					w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
				} else {
#pragma unroll
					for(int i = 0; i < 16; ++i)
						if(i < live) w[i >> 2] |= uint32_t(A.in[my + i]) << (8 * (i & 3));
				}
			}
			uint32_t prev = __shfl_up_sync(0xffffffffu, w[3] >> 24, 1);
			if(lane == 0 && my < A.n) prev = my == 0 ? A.prev0 : uint32_t(A.in[my - 1]);

			uint32_t my_bits = 0;
			if(FMT != FMT_WIDE) {
				uint32_t e[16];
				uint32_t floor = 0xffffffffu;
				if(live == 16) my_bits = gather_box<FMT, true>(A, table, w, live, prev, e, floor);
				else my_bits = gather_box<FMT, false>(A, table, w, live, prev, e, floor);
				if(floor == 0) {   // rare: count the symbols without a codeword exactly
#pragma unroll
					for(int i = 0; i < 16; ++i) dropped += (i < live && e[i] == 0) ? 1u : 0u;
				}
				// ---- 2. block exclusive scan ----
				uint32_t incl = my_bits;
#pragma unroll
				for(int d = 1; d < 32; d <<= 1) {
					const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
					if(lane >= uint32_t(d)) incl += t;
				}
				if(lane == 31) warp_sums[warp] = incl;
				__syncthreads();
				uint32_t before = 0, round_bits = 0;
#pragma unroll
				for(int k = 0; k < kEncThreads / 32; ++k) {
					const uint32_t s = warp_sums[k];
					if(uint32_t(k) < warp) before += s;
					round_bits += s;
				}
				const uint32_t pos = tile_bits + before + incl - my_bits;
				// ---- 3. pack: hi holds `fill` (< 32) pending bits left-aligned; branch-free flush ----
				uint32_t* word = stage + (pos >> 5);
				uint32_t fill = pos & 31, hi = 0;
#pragma unroll
				for(int i = 0; i < 16; ++i) {
					const uint32_t code = e[i] & ~31u, len = e[i] & 31u;
					hi |= code >> fill;
					const uint32_t spill = __funnelshift_r(0u, code, fill);   // code << (32 - fill), 0 when fill == 0
					fill += len;
					const bool full = fill >= 32;
					if(full) atomicOr(word, hi);
					word += full ? 1 : 0;
					hi = full ? spill : hi;
					fill &= 31;
				}
				if(fill) atomicOr(word, hi);
				tile_bits += round_bits;
			} else {
				// ---- u64 entries (codewords up to 56 bits) ----
				unsigned long long e[16];
#pragma unroll
				for(int i = 0; i < 16; ++i) {
					const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 255u;
					unsigned long long ent = 0;
					if(i < live) {
						ent = __ldg(A.wide + ((A.order ? prev : 0u) << 8) + c);
						if(ent == 0) ++dropped;
					}
					e[i] = ent;
					my_bits += uint32_t(ent >> 56);
					prev = c;
				}
				uint32_t incl = my_bits;
#pragma unroll
				for(int d = 1; d < 32; d <<= 1) {
					const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
					if(lane >= uint32_t(d)) incl += t;
				}
				if(lane == 31) warp_sums[warp] = incl;
				__syncthreads();
				uint32_t before = 0, round_bits = 0;
#pragma unroll
				for(int k = 0; k < kEncThreads / 32; ++k) {
					const uint32_t s = warp_sums[k];
					if(uint32_t(k) < warp) before += s;
					round_bits += s;
				}
				const uint32_t pos = tile_bits + before + incl - my_bits;
				if(my_bits) {
					uint32_t word = pos >> 5, fill = pos & 31;
					unsigned long long acc = 0;
					auto push = [&](uint32_t code, uint32_t len) {   // len <= 32, fill < 32
						acc = (acc << len) | code;
						fill += len;
						if(fill >= 32) {
							atomicOr(&stage[word], uint32_t(acc >> (fill - 32)));
							++word;
							fill -= 32;
						}
					};
#pragma unroll
					for(int i = 0; i < 16; ++i) {
						const uint32_t len = uint32_t(e[i] >> 56);
						if(len > 32) { push(uint32_t((e[i] & 0x00ffffffffffffffull) >> 32), len - 32); push(uint32_t(e[i]), 32); }
						else if(len) push(uint32_t(e[i]), len);
					}
					if(fill) atomicOr(&stage[word], uint32_t(acc) << (32 - fill));
				}
				tile_bits += round_bits;
			}
			__syncthreads();   // warp_sums reuse + staged bits visible
		}

		// ---- 4. publish + warp-wide look-back ----
		if(warp == 0) {
			const uint32_t tcount = tile_bits < 31 ? tile_bits : 31;
			const uint32_t own_tail = stage_bits(stage, tile_bits - tcount, tcount);
			unsigned long long excl_bits;
			uint32_t excl_tail;
			look_back(tile, tile_bits, own_tail, A.agg, A.inc, excl_bits, excl_tail);
			if(lane == 0) {
				s_prefix_bits = excl_bits;
				s_prefix_tail = excl_tail;
				if(tile == A.n_tiles - 1) A.result[0] = excl_bits + tile_bits;
			}
		}
		__syncthreads();

		// ---- 5. funnel-shift copy-out ----
		const unsigned long long g0 = A.bit0 + s_prefix_bits;     // global bit index of the tile's first bit
		const unsigned long long g1 = g0 + tile_bits;
		const uint32_t s = uint32_t(g0 & 31);
		const unsigned long long w0 = g0 >> 5;
		unsigned long long w1 = g1 >> 5;                          // words [w0, w1) are completed by this tile
		if(tile == A.n_tiles - 1 && (g1 & 31)) ++w1;              // the stream's last partial word, zero padded
		const uint32_t nw = uint32_t(w1 - w0);
		if(w1 > A.out_capacity_words) {
			if(tid == 0) A.result[2] = 1;                         // capacity error; this tile writes nothing
		} else {
			const uint32_t carry = s ? (s_prefix_tail & ((1u << s) - 1u)) : 0u;   // predecessor bits that open word w0
			for(uint32_t j = tid; j < nw; j += kEncThreads) {
				const uint32_t hi = j ? stage[j - 1] : carry;
				const uint32_t v = __funnelshift_r(stage[j], hi, s);
				A.out_words[w0 + j] = __byte_perm(v, 0, 0x0123);
			}
		}
		__syncthreads();
		const uint32_t used = (tile_bits + 31) / 32 + 1;
		for(uint32_t j = tid; j < used; j += kEncThreads) stage[j] = 0;
		// the barrier at the top of the loop orders this zeroing before the next tile's packing
	}
	dropped = uint32_t(warp_sum(dropped));
	if(lane == 0 && dropped) atomicAdd(A.result + 1, (unsigned long long) dropped);
}

template <int ROUNDS, int FMT>
int launch_variant(bool aligned, const EncArgs& args, size_t smem_bytes, cudaStream_t st) {
	auto kern = aligned ? encode_kernel<ROUNDS, FMT, true> : encode_kernel<ROUNDS, FMT, false>;
	MH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
	int per_sm = 0;
	MH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEncThreads, smem_bytes));
	if(per_sm < 1) per_sm = 1;
	uint64_t grid = uint64_t(sm_count()) * per_sm;
	if(grid > args.n_tiles) grid = args.n_tiles;
	{
		ProfScope p("encode_kernel", st);
		kern<<<unsigned(grid), kEncThreads, smem_bytes, st>>>(args);
	}
	count_launch(1);
	MH_CUDA(cudaGetLastError());
	return MH_OK;
}

}  // namespace

uint64_t encode_tiles_for(uint64_t n) { return (n + kEncRoundBytes - 1) / kEncRoundBytes; }

int launch_encode(const uint8_t* d_in, uint64_t n, uint8_t prev0, const mh_codebook* cb, uint64_t bit_base,
                  uint8_t* d_out, uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws, cudaStream_t st) {
	if(!cb || !cb->d_enc || !d_result || (!d_in && n) || (!d_out && n)) return MH_ERR_INVALID_ARG;
	if(reinterpret_cast<uint64_t>(d_out) & 3) return MH_ERR_INVALID_ARG;
	if(!ws || !ws->enc_desc) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(d_result, 0, 4 * sizeof(unsigned long long), st));
	if(n == 0) return MH_OK;
	// Tile size: the staged bits of one tile must fit the staging area whatever the input, so rounds x round bytes x
	// longest codeword bounds it. Prefer 2 rounds (16 KiB tiles); the staging area is sized to that bound.
	const int maxb = cb->max_bits > 0 ? cb->max_bits : 1;
	int rounds = 1;
	if(uint64_t(4) * kEncRoundBytes * maxb <= uint64_t(kEncStageMaxWords) * 32 / 2) rounds = 4;
	else if(uint64_t(2) * kEncRoundBytes * maxb <= uint64_t(kEncStageMaxWords) * 32) rounds = 2;
	const uint32_t stage_words = uint32_t((uint64_t(rounds) * kEncRoundBytes * maxb + 31) / 32 + 4);
	const uint64_t tile_bytes = uint64_t(rounds) * kEncRoundBytes;
	const uint64_t tiles = (n + tile_bytes - 1) / tile_bytes;
	if(tiles > ws->enc_tiles_cap || tiles > 0x7fffffffull) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(ws->enc_desc, 0, 2 * tiles * sizeof(uint64_t), st));   // aggregate words, then inclusive words
	MH_CUDA(cudaMemsetAsync(ws->counters, 0, sizeof(uint32_t), st));

	EncArgs a;
	a.in = d_in; a.n = n; a.prev0 = prev0; a.order = cb->order;
	a.wide = reinterpret_cast<const unsigned long long*>(cb->d_enc);
	a.box = cb->d_box; a.box_lo = cb->box_lo; a.box_r = cb->box_r;
	a.bit0 = uint32_t(bit_base & 7);
	a.stage_words = stage_words;
	a.out_words = reinterpret_cast<uint32_t*>(d_out);
	a.out_capacity_words = out_capacity / 4;
	a.agg = reinterpret_cast<unsigned long long*>(ws->enc_desc);
	a.inc = a.agg + tiles;
	a.ticket = ws->counters;
	a.result = d_result;
	a.n_tiles = uint32_t(tiles);
	const bool aligned = (reinterpret_cast<uint64_t>(d_in) & 15) == 0;
	const size_t stage_bytes = (size_t(stage_words) + 4) * sizeof(uint32_t);

	const char* fmt_env = getenv("MH_ENC_FMT");   // experiments / tests: force a table format (1: box in global, 2: wide)
	const int force_fmt = fmt_env ? atoi(fmt_env) : -1;
	int fmt = FMT_WIDE;
	size_t smem = stage_bytes;
	if(cb->has_box) {
		const size_t table_bytes = size_t(cb->order ? (cb->box_r + 1) * (cb->box_r + 1) : 256) * 4;
		if(table_bytes <= size_t(kEncBoxSmemLimit)) { fmt = FMT_BOX_SMEM; smem = stage_bytes + table_bytes; }
		else fmt = FMT_BOX_GLOBAL;
		if(force_fmt == FMT_BOX_GLOBAL) { fmt = FMT_BOX_GLOBAL; smem = stage_bytes; }
	}
	if(force_fmt == FMT_WIDE) { fmt = FMT_WIDE; smem = stage_bytes; }
	if(fmt == FMT_WIDE) {
		if(maxb > 32 || rounds == 1) return launch_variant<1, FMT_WIDE>(aligned, a, smem, st);
		if(rounds == 2) return launch_variant<2, FMT_WIDE>(aligned, a, smem, st);
		return launch_variant<4, FMT_WIDE>(aligned, a, smem, st);
	}
	if(fmt == FMT_BOX_SMEM) {
		if(rounds == 4) return launch_variant<4, FMT_BOX_SMEM>(aligned, a, smem, st);
		if(rounds == 2) return launch_variant<2, FMT_BOX_SMEM>(aligned, a, smem, st);
		return launch_variant<1, FMT_BOX_SMEM>(aligned, a, smem, st);
	}
	if(rounds == 4) return launch_variant<4, FMT_BOX_GLOBAL>(aligned, a, smem, st);
	if(rounds == 2) return launch_variant<2, FMT_BOX_GLOBAL>(aligned, a, smem, st);
	return launch_variant<1, FMT_BOX_GLOBAL>(aligned, a, smem, st);
}

}  // namespace mh
