// mh_encode.cu — kernel 2: table-lookup encoder with a single-pass decoupled look-back scan and bit packing.
//
// Replaces the loop of i_coding_provider::compress (reference src/coding.cpp:67-78) together with the bit packer
// bitbuffer::push_encoding_descriptor / push_byte / flush (src/bitbuffer.cpp:21-73, :170-180): for every input
// byte c with predecessor prev, append codeword[prev][c] MSB-first to one contiguous bit stream.
//
// Shape of the kernel. One CTA owns one tile of ROUNDS x 4 KiB of input; tiles are handed out
// in order by an atomic ticket so a tile only ever waits on tiles that are already running.
//   1. per round, each thread takes 16 consecutive bytes with one coalesced 128-bit load (the byte before them
//      comes from the neighbouring lane), gathers the 16 (length, code) entries and sums the lengths;
//   2. a block-wide exclusive scan of the per-thread bit counts gives every thread its bit offset in the tile;
//   3. each thread shifts its codewords into a 64-bit accumulator and emits whole 32-bit words into a zeroed
//      shared-memory staging area — plain stores for words it owns entirely, atomicOr only for its first and
//      last (shared) word;
//   4. the tile publishes (bit count, last 31 bits) and resolves its global bit offset by decoupled look-back
//      over the predecessors' descriptors. The scanned quantity is the pair (bits, tail) under the monoid
//      "concatenate and keep the last 31 bits", so the partially filled 32-bit word at a tile seam is completed
//      by the tile that ends it: every output word has exactly one writer and the output needs no pre-zeroing
//      and no global atomics;
//   5. the staged bits are funnel-shifted by the tile's global bit phase and written with coalesced 32-bit
//      stores, byte-swapped so that stream bit p lands in byte p/8, bit 7 - p%8 (src/bitbuffer.cpp:12).
#include "mh_internal.hpp"

namespace mh {

namespace {

struct BitsTail {
	unsigned long long bits;
	uint32_t tail;   // last min(bits, 31) bits, right-aligned, zero above
};

// concatenate a then b, keep the last 31 bits
__device__ __forceinline__ BitsTail concat(const BitsTail& a, const BitsTail& b) {
	BitsTail r;
	r.bits = a.bits + b.bits;
	r.tail = b.bits >= 31 ? b.tail : (uint32_t((uint64_t(a.tail) << b.bits) | b.tail) & 0x7fffffffu);
	return r;
}

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}

// bits [pos, pos + count) of the MSB-first word array, count <= 32, right-aligned
__device__ __forceinline__ uint32_t stage_bits(const uint32_t* stage, uint32_t pos, uint32_t count) {
	if(count == 0) return 0;
	const uint32_t w = pos >> 5, off = pos & 31;
	const uint64_t two = (uint64_t(stage[w]) << 32) | stage[w + 1];
	return uint32_t((two << off) >> (64 - count));
}

template <int ROUNDS, bool LONG_CODES, bool ALIGNED>
__global__ void __launch_bounds__(kEncThreads) encode_kernel(
    const uint8_t* __restrict__ in, uint64_t n, uint32_t prev0, const unsigned long long* __restrict__ book, int order,
    uint32_t bit0, uint32_t* __restrict__ out_words, uint64_t out_capacity_words, unsigned long long* desc,
    uint32_t* tail_agg, uint32_t* tail_inc, uint32_t* ticket, unsigned long long* result, uint32_t n_tiles) {
	__shared__ uint32_t stage[kEncStageWords + 2];
	__shared__ uint32_t warp_sums[kEncThreads / 32];
	__shared__ uint32_t s_tile;
	__shared__ unsigned long long s_prefix_bits;
	__shared__ uint32_t s_prefix_tail;
	__shared__ uint32_t s_dropped;

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if(tid == 0) { s_tile = atomicAdd(ticket, 1u); s_dropped = 0; }
	for(uint32_t i = tid; i < kEncStageWords + 2; i += kEncThreads) stage[i] = 0;
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint64_t tile_base = uint64_t(tile) * (ROUNDS * kEncRoundBytes);

	uint32_t tile_bits = 0;   // running bit count of the tile (uniform across the block)
	uint32_t dropped = 0;
#pragma unroll 1
	for(int r = 0; r < ROUNDS; ++r) {
		const uint64_t my = tile_base + uint64_t(r) * kEncRoundBytes + tid * 16;
		if(tile_base + uint64_t(r) * kEncRoundBytes >= n) break;   // uniform
		// ---- 1. load 16 bytes + the byte before them ----
		uint32_t w[4] = {0, 0, 0, 0};
		int live = 0;   // how many of my 16 bytes exist
		if(my < n) {
			live = n - my >= 16 ? 16 : int(n - my);
			if(ALIGNED && live == 16) {
				const uint4 v = ld_stream_128(in + my);
				w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
			} else {
				for(int i = 0; i < live; ++i) w[i >> 2] |= uint32_t(in[my + i]) << (8 * (i & 3));
			}
		}
		uint32_t prev = __shfl_up_sync(0xffffffffu, w[3] >> 24, 1);
		if(lane == 0 && my < n) prev = my == 0 ? prev0 : uint32_t(in[my - 1]);
		// ---- gather entries, sum lengths ----
		unsigned long long e[16];
		uint32_t my_bits = 0;
#pragma unroll
		for(int i = 0; i < 16; ++i) {
			const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 255u;
			unsigned long long ent = 0;
			if(i < live) {
				ent = __ldg(book + ((order ? prev : 0u) << 8) + c);
				if(ent == 0) ++dropped;
			}
			e[i] = ent;
			my_bits += uint32_t(ent >> 56);
			prev = c;
		}
		// ---- 2. block exclusive scan of my_bits ----
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
		uint32_t pos = tile_bits + before + incl - my_bits;
		// ---- 3. pack into the staging area ----
		if(my_bits) {
			uint32_t word = pos >> 5;
			uint32_t fill = pos & 31;
			unsigned long long acc = 0;
			bool first = true;
			auto push = [&](uint32_t code, uint32_t len) {   // len <= 32, fill < 32
				acc = (acc << len) | code;
				fill += len;
				if(fill >= 32) {
					const uint32_t full = uint32_t(acc >> (fill - 32));
					if(first) { atomicOr(&stage[word], full); first = false; }
					else stage[word] = full;
					++word;
					fill -= 32;
				}
			};
#pragma unroll
			for(int i = 0; i < 16; ++i) {
				const uint32_t len = uint32_t(e[i] >> 56);
				if(LONG_CODES) {
					if(len > 32) { push(uint32_t((e[i] & 0x00ffffffffffffffull) >> 32), len - 32); push(uint32_t(e[i]), 32); }
					else if(len) push(uint32_t(e[i]), len);
				} else {
					push(uint32_t(e[i]), len);   // len == 0 is a no-op
				}
			}
			if(fill) atomicOr(&stage[word], uint32_t(acc) << (32 - fill));
		}
		tile_bits += round_bits;
		__syncthreads();   // warp_sums reuse + staging visible
	}
	if(dropped) atomicAdd(&s_dropped, dropped);

	// ---- 4. publish + look back (thread 0), others wait ----
	if(tid == 0) {
		BitsTail own;
		own.bits = tile_bits;
		const uint32_t tcount = tile_bits < 31 ? tile_bits : 31;
		own.tail = stage_bits(stage, tile_bits - tcount, tcount);
		BitsTail excl = {0ull, 0u};
		if(tile > 0) {
			tail_agg[tile] = own.tail;
			st_release(desc + tile, kDescAggregate | own.bits);
			BitsTail run = {0ull, 0u};
			bool have = false;
			for(int64_t j = int64_t(tile) - 1; j >= 0; --j) {
				unsigned long long d;
				do { d = ld_acquire(desc + j); } while((d >> 62) == 0);
				BitsTail t;
				t.bits = d & kDescValueMask;
				const bool inclusive = (d >> 62) == 2;
				t.tail = *(volatile uint32_t*) ((inclusive ? tail_inc : tail_agg) + j);
				run = have ? concat(t, run) : t;
				have = true;
				if(inclusive) break;
			}
			excl = run;
		}
		const BitsTail incl = concat(excl, own);
		tail_inc[tile] = incl.tail;
		st_release(desc + tile, kDescInclusive | incl.bits);
		s_prefix_bits = excl.bits;
		s_prefix_tail = excl.tail;
		if(tile == n_tiles - 1) result[0] = incl.bits;
	}
	__syncthreads();
	if(tid == 0 && s_dropped) atomicAdd(result + 1, (unsigned long long) s_dropped);

	// ---- 5. funnel-shift copy-out ----
	const unsigned long long g0 = bit0 + s_prefix_bits;       // global bit index of the tile's first bit
	const unsigned long long g1 = g0 + tile_bits;
	const uint32_t s = uint32_t(g0 & 31);
	const unsigned long long w0 = g0 >> 5;
	unsigned long long w1 = g1 >> 5;                          // words [w0, w1) are completed by this tile
	if(tile == n_tiles - 1 && (g1 & 31)) ++w1;                // the stream's last partial word, zero padded
	const uint32_t nw = uint32_t(w1 - w0);
	if(w1 > out_capacity_words) {
		if(tid == 0) result[2] = 1;                           // capacity error flag; nothing is written
		return;
	}
	const uint32_t carry = s ? (s_prefix_tail & ((1u << s) - 1u)) : 0u;   // predecessor bits that open word w0
	for(uint32_t j = tid; j < nw; j += kEncThreads) {
		const uint32_t hi = j ? stage[j - 1] : carry;
		const uint32_t v = __funnelshift_r(stage[j], hi, s);
		out_words[w0 + j] = __byte_perm(v, 0, 0x0123);
	}
}

}  // namespace

uint64_t encode_tiles_for(uint64_t n) { return (n + kEncRoundBytes - 1) / kEncRoundBytes; }

namespace {

template <int ROUNDS, bool LONG_CODES>
int launch_variant(bool aligned, uint32_t n_tiles, cudaStream_t st, const uint8_t* d_in, uint64_t n, uint32_t prev0,
                   const mh_codebook* cb, uint32_t bit0, uint32_t* out_words, uint64_t cap_words, mh_workspace* ws,
                   unsigned long long* d_result) {
	ProfScope p("encode_kernel", st);
	if(aligned)
		encode_kernel<ROUNDS, LONG_CODES, true><<<n_tiles, kEncThreads, 0, st>>>(d_in, n, prev0, (const unsigned long long*) cb->d_enc, cb->order, bit0,
		    out_words, cap_words, (unsigned long long*) ws->enc_desc, ws->enc_tail_agg, ws->enc_tail_inc, ws->counters, d_result, n_tiles);
	else
		encode_kernel<ROUNDS, LONG_CODES, false><<<n_tiles, kEncThreads, 0, st>>>(d_in, n, prev0, (const unsigned long long*) cb->d_enc, cb->order, bit0,
		    out_words, cap_words, (unsigned long long*) ws->enc_desc, ws->enc_tail_agg, ws->enc_tail_inc, ws->counters, d_result, n_tiles);
	count_launch(1);
	MH_CUDA(cudaGetLastError());
	return MH_OK;
}

}  // namespace

int launch_encode(const uint8_t* d_in, uint64_t n, uint8_t prev0, const mh_codebook* cb, uint64_t bit_base,
                  uint8_t* d_out, uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws, cudaStream_t st) {
	if(!cb || !cb->d_enc || !d_result || (!d_in && n) || (!d_out && n)) return MH_ERR_INVALID_ARG;
	if(reinterpret_cast<uint64_t>(d_out) & 3) return MH_ERR_INVALID_ARG;
	if(!ws || !ws->enc_desc) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(d_result, 0, 4 * sizeof(unsigned long long), st));
	if(n == 0) return MH_OK;
	// tile size: the staged bits of one tile must fit kEncStageWords words whatever the input
	const int maxb = cb->max_bits > 0 ? cb->max_bits : 1;
	const uint64_t stage_bits_cap = uint64_t(kEncStageWords) * 32;
	int rounds = 1;
	if(uint64_t(4) * kEncRoundBytes * maxb <= stage_bits_cap) rounds = 4;
	else if(uint64_t(2) * kEncRoundBytes * maxb <= stage_bits_cap) rounds = 2;
	const uint64_t tile_bytes = uint64_t(rounds) * kEncRoundBytes;
	const uint64_t tiles = (n + tile_bytes - 1) / tile_bytes;
	if(tiles > ws->enc_tiles_cap || tiles > 0x7fffffffull) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(ws->enc_desc, 0, tiles * sizeof(uint64_t), st));
	MH_CUDA(cudaMemsetAsync(ws->counters, 0, sizeof(uint32_t), st));
	const bool aligned = (reinterpret_cast<uint64_t>(d_in) & 15) == 0;
	const bool long_codes = maxb > 32;
	const uint32_t bit0 = uint32_t(bit_base & 7);
	uint32_t* out_words = reinterpret_cast<uint32_t*>(d_out);
	const uint64_t cap_words = out_capacity / 4;
	const uint32_t nt = uint32_t(tiles);
	if(long_codes) return launch_variant<1, true>(aligned, nt, st, d_in, n, prev0, cb, bit0, out_words, cap_words, ws, d_result);
	if(rounds == 4) return launch_variant<4, false>(aligned, nt, st, d_in, n, prev0, cb, bit0, out_words, cap_words, ws, d_result);
	if(rounds == 2) return launch_variant<2, false>(aligned, nt, st, d_in, n, prev0, cb, bit0, out_words, cap_words, ws, d_result);
	return launch_variant<1, false>(aligned, nt, st, d_in, n, prev0, cb, bit0, out_words, cap_words, ws, d_result);
}

}  // namespace mh
