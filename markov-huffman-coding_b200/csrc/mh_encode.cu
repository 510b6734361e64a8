// mh_encode.cu — kernel 2: table-lookup encoder with a single-pass decoupled look-back scan and bit packing.
//
// Replaces the loop of i_coding_provider::compress (reference src/coding.cpp:67-78) together with the bit packer
// bitbuffer::push_encoding_descriptor / push_byte / flush (src/bitbuffer.cpp:21-73, :170-180): for every input
// byte c with predecessor prev, append codeword[prev][c] MSB-first to one contiguous bit stream.
//
// Shape of the kernel. Persistent CTAs (two per SM) of 15 worker warps and one scanner warp pull
// tiles of 480 x SPT input bytes (SPT = 32 or 16 symbols per thread) from an atomic ticket, so a tile only ever
// waits on tiles that are already running.
//   0. once per CTA the codebook is staged in shared memory: for text, context rows (one 256-entry u32 row per live
//      context, entry = 5-bit length | next row | 16-bit code, the rare longer codeword escapes to the wide table);
//      otherwise u32 entries (5-bit length | 27-bit code) for the RxR box of byte values the table uses, with a zero
//      border that bytes outside the box are clamped onto; tables that fit neither are gathered from global memory;
//   1. each worker thread takes SPT consecutive bytes with coalesced 128-bit loads (the byte before them comes from
//      the neighbouring lane by shuffle), looks up its entries and sums their lengths;
//   2. a block-wide exclusive scan of the per-thread bit counts gives every thread its bit offset in the tile; the
//      tile's bit count is PUBLISHED RIGHT AWAY, so successors can resolve their offsets while this tile packs;
//   3. codewords are merged pairwise, then by fours, in registers and funnelled through a bit window; only completed
//      32-bit words are ORed into the zeroed shared-memory staging area (neighbouring threads share their boundary
//      words, hence the OR); thread 0 publishes the tile's last 31 bits;
//   4. the scanner warp resolves the tile's global bit offset by a warp-wide decoupled look-back (32 predecessors per
//      poll, relaxed loads that bypass L1) as soon as the bit count is known — one iteration before the workers need
//      it — and hands it over through a named barrier. It also writes the one output word the tile shares with its
//      predecessor (the predecessor's tail bits, then this tile's first bits), so every output word has exactly one
//      writer: no pre-zeroed output, no global atomics, no second pass;
//   5. one iteration later the workers funnel-shift the staged bits by the tile's global bit phase and write them
//      with coalesced 32-bit stores, byte-swapped so that stream bit p lands in byte p/8, bit 7 - p%8
//      (src/bitbuffer.cpp:12); the words just copied are re-zeroed for the next tile.
// Device-built tables (the compress path of sessions, shards and the bench) are encoded by TWO launches of this kernel: an
// optimistic instance with SPT = 64 (a third of the 32-symbol kernel's instructions were paid per tile, not per symbol;
// its quads are one word + a packed length, its staging area is what shared memory leaves, and it declines tables and
// fails on tiles it is not made for) and the ordinary SPT = 32 instance queued behind it, which reads the first one's
// flag and either hands its results over and leaves, or encodes the stream itself (launch_encode).
#include "mh_internal.hpp"

namespace mh {

namespace {

constexpr int FMT_BOX_SMEM = 0;     // u32 box table in shared memory
constexpr int FMT_BOX_GLOBAL = 1;   // u32 box table gathered from global memory
constexpr int FMT_WIDE = 2;         // u64 entries (8-bit length | 56-bit code) gathered from global memory
constexpr int FMT_CTX = 3;          // u32 context rows in shared memory: len << 27 | next row << 16 | code (<= 16 bits)

constexpr uint64_t kAgg = 1ull << 62;   // aggregate word : kAgg | bits of this tile
constexpr uint64_t kInc = 2ull << 62;   // inclusive word : kInc | bits of all tiles up to and including this one
constexpr uint64_t kLow31 = 0x7fffffffull;
constexpr uint32_t kTailValid = 0x80000000u;   // tail word: kTailValid | last min(bits, 31) bits of this tile

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed32(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed32(uint32_t* p, uint32_t v) {
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
// ---- bulk copy (TMA) of a tile's input into shared memory, completion through an mbarrier -----------------------
// One thread arms the barrier with the byte count and issues ONE cp.async.bulk for the whole tile (SASS: UBLKCP +
// SYNCS): no worker issues a global load for its input, and the copy of the next tile runs while this one is packed.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load_tile(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "WAIT_%=:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@!p bra WAIT_%=;\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
	uint4 r;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
	return r;
}

// Shared memory through explicit 32-bit shared-space addresses: keeps generic-address arithmetic out of the hot
// loops.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void red_or_if(uint32_t addr, uint32_t v, bool go) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"(uint32_t(go)) : "memory");
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

struct EncArgs {
	const uint8_t* in;
	uint64_t n;
	uint32_t prev0;
	int order;
	const unsigned long long* wide;   // FMT_WIDE table: len << 56 | code
	const uint32_t* box;              // u32 box table with a zero border: (R + 1) x (R + 1), or 256 entries (order 0)
	uint32_t box_lo, box_r;
	const uint32_t* ctx;              // FMT_CTX table: ctx_rows x 256 entries, the last row is the null row
	uint32_t ctx_rows;
	uint32_t bit0;
	const unsigned long long* bit_base_dev;   // optional: the global bit offset lives in device memory (its low 3 bits replace bit0)
	const unsigned long long* prev0_dev;      // optional: the byte before in[0] lives in device memory
	const uint32_t* meta;                     // optional: tables built on the device: rows / status / longest codeword live there
	uint32_t launched_bits;                   // ... and this launch was sized for at most launched_bits-bit codewords (tile size) ...
	uint32_t smem_bytes;                      // ... and smem_bytes of dynamic shared memory (table + staging area + input tile)
	uint32_t stage_words;
	uint32_t* out_words;
	uint64_t out_capacity_words;
	unsigned long long* agg;
	unsigned long long* inc;
	unsigned long long* grp_acc;   // warp-private tiles: per group of tiles, tiles counted << 48 | bits (one per 128-byte line)
	unsigned long long* grp_inc;   // ... and valid << 63 | bits up to and including the group
	uint32_t* tail;
	uint32_t* ticket;
	unsigned long long* result;
	uint32_t n_tiles;
	uint32_t full_tiles;   // tiles that are complete (n / tile bytes): computed on the host — in the kernel the 64-bit division was redone every tile
	uint32_t* fail;                           // SPT = 64 (optimistic launch): raised when the launch declines the tables or a tile's bits do not fit its staging area
	const uint32_t* gate;                     // launched behind the SPT = 64 variant: run only if *gate != 0, else hand its results over and leave
	const unsigned long long* alt_result;     // ... the results of the SPT = 64 launch (bits, dropped symbols, capacity error)
};

// Warp-wide decoupled look-back over a window of kWin x 32 predecessors per poll (one warp-load measured best). All 32 lanes of warp 0 call this.
// Returns, on every lane, the number of bits that precede this tile and the bit count of tile - 1; publishes the
// tile's inclusive prefix.
constexpr int kWin = 1;

__device__ __forceinline__ void look_back(uint32_t tile, uint32_t own_bits, const EncArgs& A, unsigned long long& excl_bits,
                                          uint32_t& nearest_bits) {
	const uint32_t lane = threadIdx.x & 31;
	unsigned long long sum = 0;
	unsigned long long nearest = 0;
	bool first_window = true;
	long long base = (long long) tile - 1;
	while(base >= 0) {
		unsigned long long a[kWin], p[kWin];
#pragma unroll
		for(int k = 0; k < kWin; ++k) {
			const long long j = base - (k * 32 + int(lane));
			a[k] = kAgg; p[k] = kInc;   // tiles before the stream: inclusive, zero bits
			if(j >= 0) { p[k] = ld_relaxed(A.inc + j); a[k] = ld_relaxed(A.agg + j); }
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
				if(k == 0 && first_window) need |= 1u;   // tile - 1's own bit count is needed for the seam word
				if((has_agg & need) != need) ready = false;
			}
		}
		if(!ready) continue;   // a predecessor has not published yet: poll again
		if(first_window) { nearest = __shfl_sync(0xffffffffu, a[0], 0) & kDescValueMask; first_window = false; }
		unsigned long long v = 0;
#pragma unroll
		for(int k = 0; k < kWin; ++k) {
			const int idx = k * 32 + int(lane);
			if(stop < 0 || idx < stop) v += a[k] & kDescValueMask;
			else if(idx == stop) v += p[k] & kDescValueMask;
		}
		sum += warp_sum(v);
		if(stop >= 0) break;
		base -= kWin * 32;
	}
	excl_bits = sum;
	nearest_bits = nearest > 31 ? 31u : uint32_t(nearest);
	if(lane == 0) st_relaxed(A.inc + tile, kInc | (sum + own_bits));
}

// The last min(bits before this tile, 31) bits of the stream before this tile. Lane 0 of warp 0 only.
__device__ __forceinline__ uint32_t predecessor_tail(uint32_t tile, uint32_t nearest_bits, const EncArgs& A) {
	if(tile == 0) return 0;
	uint32_t tv;
	do { tv = ld_relaxed32(A.tail + tile - 1); } while(!(tv & kTailValid));
	uint32_t t_tail = tv & uint32_t(kLow31), t_bits = nearest_bits;
	// Rare (a predecessor produced fewer than 31 bits, e.g. dropped symbols): keep prepending older tiles.
	for(long long j = (long long) tile - 2; j >= 0 && t_bits < 31; --j) {
		unsigned long long a;
		do { a = ld_relaxed(A.agg + j); } while((a >> 62) != 1);
		do { tv = ld_relaxed32(A.tail + j); } while(!(tv & kTailValid));
		const unsigned long long b = a & kDescValueMask;
		t_tail = uint32_t(((uint64_t(tv & uint32_t(kLow31)) << t_bits) | t_tail) & kLow31);
		t_bits = b + t_bits > 31 ? 31u : uint32_t(b) + t_bits;
	}
	return t_tail;
}

// Named barriers. 0: the whole CTA (__syncthreads); 1: the worker warps; 2, 3: "tile index and bit count of iteration i
// are in shared memory" (workers arrive, the scanner warp waits); 4, 5: "the prefix of the tile packed in iteration i is
// in shared memory" (the scanner arrives, the workers wait). Arrive / sync pairs order the shared-memory hand-over.
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(kEncThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kEncCtaThreads) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEncCtaThreads) : "memory"); }

// Block-wide exclusive scan over the worker threads (one value per thread). One barrier; `total` is the block sum.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
	constexpr int NW = kEncThreads / 32;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
		if(lane >= uint32_t(d)) incl += t;
	}
	if(lane == 31) warp_sums[warp] = incl;
	bar_workers();
	const uint32_t ws = lane < NW ? warp_sums[lane] : 0u;   // every warp scans the warp totals itself
	uint32_t wincl = ws;
#pragma unroll
	for(int d = 1; d < 16; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, d);
		if(lane >= uint32_t(d)) wincl += t;
	}
	total = __shfl_sync(0xffffffffu, wincl, NW - 1);
	const uint32_t before = __shfl_sync(0xffffffffu, wincl - ws, warp);
	return before + incl - v;
}

// A thread's bit window: `fill` (< 32) pending bits left-aligned in `hi`; completed words are ORed into the stage.
struct Packer {
	uint32_t word;   // shared-space byte address of the word being filled
	uint32_t hi, fill;
	__device__ __forceinline__ void start(uint32_t stage_addr, uint32_t pos) { word = stage_addr + ((pos >> 5) << 2); hi = 0; fill = pos & 31; }
	// append a right-aligned unit of len <= 64 bits (len == 0: v == 0)
	__device__ __forceinline__ void put(unsigned long long v, uint32_t len) {
		const uint32_t vh = uint32_t(v >> 32), vl = uint32_t(v);
		const uint32_t up = (64u - len) & 63u;                      // left-align: q = v << up (len == 64 or 0 -> no shift)
		const uint32_t qh = up >= 32 ? vl << (up - 32) : __funnelshift_l(vl, vh, up);
		const uint32_t ql = up >= 32 ? 0u : vl << up;
		const uint32_t w0 = hi | (qh >> fill);
		const uint32_t w1 = __funnelshift_r(ql, qh, fill);
		const uint32_t w2 = __funnelshift_r(0u, ql, fill);          // ql << (32 - fill), 0 when fill == 0
		const uint32_t nf = fill + len;                             // < 96
		const uint32_t full = nf >> 5;                              // 0, 1 or 2 words completed
		red_or_if(word, w0, full >= 1);
		red_or_if(word + 4, w1, full == 2);
		word += full << 2;
		hi = full == 0 ? w0 : (full == 1 ? w1 : w2);
		fill = nf & 31;
	}
	__device__ __forceinline__ void finish() { red_or_if(word, hi, fill != 0); }
};

// FMT_CTX packer: `fill` (< 32) pending bits RIGHT-aligned in `pend`; whatever sits above them is garbage that never
// reaches an emitted word. Appends right-aligned units of len <= 32 bits.
struct Packer32 {
	uint32_t word;   // shared-space byte address of the word being filled
	uint32_t pend, fill;
	__device__ __forceinline__ void start(uint32_t stage_addr, uint32_t pos) { word = stage_addr + ((pos >> 5) << 2); pend = 0; fill = pos & 31; }
	__device__ __forceinline__ void put(uint32_t v, uint32_t len) {
		uint32_t tlo;
		asm("shl.b32 %0, %1, %2;" : "=r"(tlo) : "r"(pend), "r"(len));   // PTX shifts clamp: len == 32 -> 0
		tlo |= v;
		const uint32_t thi = __funnelshift_lc(pend, 0u, len);           // pend >> (32 - len), 0 for len == 0
		const uint32_t nf = fill + len;                                 // < 64
		red_or_if(word, __funnelshift_r(tlo, thi, nf), nf >= 32);       // bits [nf - 32, nf) of thi:tlo (shift wraps to nf - 32)
		word += (nf >> 5) << 2;
		pend = tlo;
		fill = nf & 31;
	}
	__device__ __forceinline__ void finish() {
		uint32_t v;
		asm("shl.b32 %0, %1, %2;" : "=r"(v) : "r"(pend), "r"(32u - fill));
		red_or_if(word, v, fill != 0);
	}
};

// Context rows sit kCtxPitch bytes apart in shared memory, not 1024: with 256-word rows the bank of (row, c) is c mod 32,
// and 32 lanes that look the SAME frequent byte up in DIFFERENT contexts — the common case in text — hit one bank
// (4.8 wavefronts per warp lookup on the bench text); five words of padding rotate every row by 5 banks: 3.1.
constexpr uint32_t kCtxPitch = (256 + 5) * 4;
// Order 0 (-h) has ONE row, and 32 lanes that look 32 different bytes up in it collide like any 32 random banks (3.5
// wavefronts per warp lookup). There the row is replicated per lane — entry c of lane l at word c * 32 + l, bank l — and
// every lookup is one wavefront (32 KiB instead of 1 KiB).
constexpr uint32_t kCtxOrder0Bytes = 256 * 32 * 4;
__host__ __device__ __forceinline__ uint32_t ctx_table_bytes(uint32_t ctx_rows, int order) { return order ? ctx_rows * kCtxPitch : kCtxOrder0Bytes; }
constexpr uint32_t kEncInbufLead = 16;   // the 16 bytes in front of a tile travel with it: the byte before the tile's first is its context

template <int SPT, int FMT, bool ALIGNED, int ORDER = 1, bool TMA = false>
__global__ void __launch_bounds__(kEncCtaThreads, FMT == FMT_CTX ? 2 : 1) encode_kernel(const EncArgs A) {   // context rows: two CTAs per SM (<= 64 registers)
	constexpr int NWORDS = SPT / 4;
	constexpr uint32_t kTileBytes = kEncThreads * SPT;
	__shared__ __align__(8) unsigned long long s_mbar;
	__shared__ uint8_t s_rank[256];   // FMT_CTX: row of every byte value as a context (what the null row's entries name)
	extern __shared__ uint32_t smem[];
	uint32_t* table = smem;   // FMT_BOX_SMEM: [(R + 1)^2] or [256]
	__shared__ uint32_t warp_sums[kEncThreads / 32];
	__shared__ uint32_t s_tile[2];
	__shared__ volatile uint32_t s_pub_tile[2], s_pub_bits[2];    // workers -> scanner: tile and bit count of iteration i
	__shared__ volatile unsigned long long s_prefix_bits[2];      // scanner -> workers: bits before that tile ...
	__shared__ volatile uint32_t s_head[2];                       // workers -> scanner: the first staged word of that tile

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t bit0 = A.bit_base_dev ? uint32_t(__ldg(A.bit_base_dev) & 7ull) : A.bit0;
	const uint32_t first_prev = A.prev0_dev ? uint32_t(__ldg(A.prev0_dev) & 255ull) : A.prev0;
	uint32_t ctx_rows = A.ctx_rows, stage_words = A.stage_words;
	if(A.gate && ld_relaxed32(A.gate) == 0) {   // the optimistic launch before this one did the work: its results are the call's results
		if(blockIdx.x == 0 && tid < 3) A.result[tid] = A.alt_result[tid];
		return;
	}
	if constexpr(SPT == 64) {
		// The optimistic variant (device-built tables only): 64 symbols per thread halve what a tile costs outside the lookups
		// and the packing (block scan, barriers, tickets, copy-out set-up: a third of the instructions at 32 symbols per thread).
		// Its staging area is what the launch's shared memory leaves next to the table and the 30 KiB input tile — not the
		// worst case of 64 x 480 longest codewords — so it declines tables it is not made for (mean codeword above 5.5 bits:
		// quads longer than 32 bits become common and they take the slow escape path; a staging area below 8 bits per symbol)
		// and raises *fail when a tile's bits do not fit after all; the SPT = 32 launch queued behind it then does the work.
		ctx_rows = __ldg(A.meta);
		const uint32_t status = __ldg(A.meta + 1), longest = __ldg(A.meta + 2);
		const unsigned long long sum_bits = __ldg(reinterpret_cast<const unsigned long long*>(A.meta + 4)), sum_count = __ldg(reinterpret_cast<const unsigned long long*>(A.meta + 6));
		const uint32_t worst = (kEncThreads * 2u * (longest ? longest : 1u) + 7u) & ~3u;
		const uint32_t fixed = ((ctx_table_bytes(ctx_rows, ORDER) + 15u) & ~15u) + kTileBytes + kEncInbufLead + 16u + 8u * 4u;
		const uint32_t room = A.smem_bytes > fixed ? ((A.smem_bytes - fixed) / 4u) & ~3u : 0u;
		stage_words = worst < room ? worst : room;
		if(status != 0 || ctx_rows > uint32_t(kEncCtxMaxRows) || longest > A.launched_bits || (stage_words < worst && stage_words < kEncThreads * 16u + 8u) ||
		   sum_bits * 2ull > sum_count * 11ull) {
			if(blockIdx.x == 0 && tid == 0) *A.fail = 1u;
			return;
		}
	} else
	if(FMT == FMT_CTX && A.meta) {   // device-built tables: check that this launch's shared memory and tile size fit them
		ctx_rows = __ldg(A.meta);
		const uint32_t status = __ldg(A.meta + 1), longest = __ldg(A.meta + 2);
		stage_words = (kEncThreads * (SPT / 32) * (longest ? longest : 1u) + 7u) & ~3u;   // as launch_encode sizes it for host-built tables
		const uint32_t need = ((ctx_table_bytes(ctx_rows, ORDER) + 15u) & ~15u) + (stage_words + 8u) * 4u + (TMA ? kTileBytes + kEncInbufLead + 16u : 0u);
		if(status != 0 || ctx_rows > uint32_t(kEncCtxMaxRows) || longest > A.launched_bits || need > A.smem_bytes) {
			if(blockIdx.x == 0 && tid == 0) A.result[3] = status ? (unsigned long long) (long long) (int) status : 1ull;   // the caller takes the host-built path
			return;
		}
	}
	uint32_t table_entries = 0;
	if(FMT == FMT_BOX_SMEM) {
		table_entries = A.order ? (A.box_r + 1) * (A.box_r + 1) : 256u;
		for(uint32_t i = tid; i < table_entries; i += kEncCtaThreads) table[i] = __ldg(A.box + i);
	}
	if(FMT == FMT_CTX) {
		table_entries = ctx_table_bytes(ctx_rows, ORDER) / 4;
		if(ORDER) for(uint32_t i = tid; i < ctx_rows * 256u; i += kEncCtaThreads) table[(i >> 8) * (kCtxPitch / 4) + (i & 255u)] = __ldg(A.ctx + i);
		else for(uint32_t i = tid; i < 256u * 32u; i += kEncCtaThreads) table[i] = __ldg(A.ctx + (i >> 5));   // the one row, once per lane
		// A quad restarts its lookup chain from the byte before it: that byte's row comes from this byte map — four rows to a
		// word, so the 32 lanes of a warp touch a handful of words (the null row's u32 entries cost ~3 wavefronts a lookup)
		if(tid < 256) s_rank[tid] = uint8_t(__ldg(A.ctx + (ctx_rows - 1) * 256u + tid) >> 16);
	}
	uint32_t* stage = smem + ((table_entries + 3) & ~3u) + 4;   // [stage_words + 4], after four zero words: stage[-1] reads as "no bits"
	const uint32_t table_sa = uint32_t(__cvta_generic_to_shared(table));
	const uint32_t stage_sa = uint32_t(__cvta_generic_to_shared(stage));
	for(uint32_t i = tid; i < stage_words + 8; i += kEncCtaThreads) stage[int(i) - 4] = 0;
	if(tid == 0) s_tile[0] = atomicAdd(A.ticket, 1u);
	// input tile buffer (TMA): behind the staging area, 16-byte aligned; tiles that are complete travel by bulk copy
	const uint32_t inbuf_sa = (stage_sa + (stage_words + 4u) * 4u + 15u) & ~15u;
	const uint32_t mbar_sa = uint32_t(__cvta_generic_to_shared(&s_mbar));
	const uint32_t full_tiles = A.full_tiles;
	auto issue_tile = [&](uint32_t tile) {   // one thread: the tile's bytes and the 16 before them (its first context)
		if(tile >= full_tiles) return;
		const uint64_t at = uint64_t(tile) * kTileBytes;
		if(tile) bulk_load_tile(inbuf_sa, A.in + at - kEncInbufLead, kTileBytes + kEncInbufLead, mbar_sa);
		else bulk_load_tile(inbuf_sa + kEncInbufLead, A.in, kTileBytes, mbar_sa);
	};
	if(TMA && tid == 0) {
		mbar_init(mbar_sa, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if(TMA && tid == kEncThreads) issue_tile(s_tile[0]);
	uint32_t in_parity = 0;

	// ---- the scanner warp: resolves every tile's global bit offset as soon as its bit count is known, one iteration
	// before the workers need it, so that nobody ever waits for the chained scan ----
	if(tid >= kEncThreads) {
		// The output word a tile shares with its predecessor (the predecessor's last bits, then this tile's first bits)
		// is written by the scanner, one iteration later: it is the only thing that needs the predecessor's tail, which
		// is published when the predecessor has been PACKED — waiting for that in the workers kept every CTA in step
		// with the slowest tile before it.
		bool have_prev = false;
		uint32_t q_tile = 0, q_bits = 0, q_near = 0;
		unsigned long long q_excl = 0;
		auto boundary_word = [&](uint32_t head) {   // lane 0: the first output word of tile q_tile, if it shares it
			const unsigned long long g0 = bit0 + q_excl, g1 = g0 + q_bits;
			const uint32_t s = uint32_t(g0 & 31);
			const unsigned long long w0 = g0 >> 5;
			unsigned long long w1 = g1 >> 5;
			if(q_tile == A.n_tiles - 1 && (g1 & 31)) ++w1;
			if(s == 0 || w1 == w0 || w1 > A.out_capacity_words) return;   // the workers' word / nothing completed / capacity error
			const uint32_t carry = predecessor_tail(q_tile, q_near, A) & ((1u << s) - 1u);
			A.out_words[w0] = __byte_perm(__funnelshift_r(head, carry, s), 0, 0x0123);
		};
		for(uint32_t it = 0;; ++it) {
			bar_wait(2 + (it & 1));
			// the workers have taken this tile's bytes out of the input buffer and the next ticket is known: its bytes travel
			// while this tile is scanned, copied out and packed
			if(TMA && lane == 0) issue_tile(s_tile[(it + 1) & 1]);
			const uint32_t tile = s_pub_tile[it & 1];
			const uint32_t head = s_head[(it - 1) & 1];   // of the tile packed in iteration it - 1 (read before the workers can reuse the slot)
			if(tile < A.n_tiles) {
				const uint32_t bits = s_pub_bits[it & 1];
				unsigned long long excl_bits;
				uint32_t nearest_bits;
				look_back(tile, bits, A, excl_bits, nearest_bits);
				if(lane == 0) {
					s_prefix_bits[it & 1] = excl_bits;
					if(tile == A.n_tiles - 1) A.result[0] = excl_bits + bits;
					__threadfence_block();
				}
				__syncwarp();
				bar_arrive(4 + (it & 1));
				if(lane == 0 && have_prev) boundary_word(head);
				__syncwarp();   // the warp meets the next barrier converged
				have_prev = true; q_tile = tile; q_bits = bits; q_near = nearest_bits; q_excl = excl_bits;
			} else {
				if(lane == 0 && have_prev) boundary_word(head);
				break;
			}
		}
		return;
	}
	const uint32_t R = A.box_r, lo = A.box_lo, pitch = R + 1;
	uint32_t dropped = 0;

	// Software pipeline: the copy-out of tile k is deferred until tile k+1 has been loaded, looked up and counted.
	// The look-back of tile k therefore starts a good part of a tile-time after its predecessors were handed out,
	// which absorbs the jitter between CTAs (every tile has to wait for the slowest tile before it), and the next
	// tile's global loads are already in flight while warp 0 polls. The staging area is reused: tile k+1 is packed
	// only after tile k has left it.
	bool pending = false;          // a packed tile is waiting in the staging area
	uint32_t p_tile = 0, p_bits = 0;
	for(uint32_t it = 0;; ++it) {
		// no barrier here: every iteration ends with one (after the pack, or after re-zeroing the staging area), and the
		// ticket read below was written before it
		const uint32_t tile = s_tile[it & 1];
		const bool valid = tile < A.n_tiles;
		if(!valid && !pending) {
			if(it == 0) {   // this CTA never got a tile (another one took two tickets first): release the scanner
				if(tid == 0) s_pub_tile[0] = 0xffffffffu;
				bar_arrive(2);
			}
			break;
		}
		if(valid && tid == 0) s_tile[(it + 1) & 1] = atomicAdd(A.ticket, 1u);   // next ticket, off the critical path
		else if(!valid && tid == 0) s_tile[(it + 1) & 1] = tile;
		const uint64_t my = uint64_t(tile) * (kEncThreads * SPT) + uint64_t(tid) * SPT;

		// ================= phase A: load, look up, count (tile `tile`) =================
		uint32_t tile_bits = 0, pos = 0;
		uint32_t e32[(FMT == FMT_BOX_SMEM || FMT == FMT_BOX_GLOBAL) ? SPT : 1];
		unsigned long long e64[FMT == FMT_WIDE ? SPT : 1];
		uint32_t q_hi[FMT == FMT_CTX && SPT != 64 ? SPT / 4 : 1], q_lo[FMT == FMT_CTX ? SPT / 4 : 1], q_len[FMT == FMT_CTX && SPT != 64 ? SPT / 4 : 1];
		uint32_t qlen4[SPT == 64 ? 4 : 1];   // SPT = 64: the quads' lengths, four to a word (a quad longer than 32 bits is escaped, so there is no q_hi)
		uint32_t esc_mask = 0;   // FMT_CTX: quads of this thread that hold a codeword longer than 16 bits
		bool overflow = false;   // SPT = 64: this tile's bits do not fit the staging area (the launch fails over to SPT = 32)
		if(valid) {
			uint32_t my_bits = 0;
			if constexpr(SPT == 64) {
				// The tile is in the input buffer (bulk copy; the stream's ragged last tile is put there by the workers, zero
				// padded): 4 x 16 bytes per thread from shared memory, one chunk at a time (registers).
				const bool from_smem = tile < full_tiles;
				const uint32_t my_sa = inbuf_sa + kEncInbufLead + tid * 64u;
				int live = 0;
				uint32_t prev = first_prev;
				if(from_smem) {
					mbar_wait(mbar_sa, in_parity);
					in_parity ^= 1u;
					live = 64;
					if(my) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(prev) : "r"(my_sa - 1u) : "memory");
				} else if(my < A.n) {
					live = A.n - my >= 64ull ? 64 : int(A.n - my);
#pragma unroll 1
					for(int k = 0; k < 16; ++k) {
						uint32_t x = 0;
#pragma unroll
						for(int b = 0; b < 4; ++b)
							if(4 * k + b < live) x |= uint32_t(A.in[my + 4 * k + b]) << (8 * b);
						asm volatile("st.shared.u32 [%0], %1;" ::"r"(my_sa + 4u * k), "r"(x) : "memory");
					}
					if(my) prev = uint32_t(A.in[my - 1]);
				}
				const uint32_t rank_sa = uint32_t(__cvta_generic_to_shared(s_rank));
				auto row_of = [&](uint32_t byte) -> uint32_t {
					uint32_t r;
					asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(rank_sa + byte));
					return table_sa + r * kCtxPitch;
				};
				uint32_t row = table_sa + lane * 4;
				if(ORDER) row = row_of(prev);
				uint32_t floor = 0xffffffffu, ceil = 0;
				auto quads64 = [&](bool checked) {
					uint32_t lastc = prev;   // the byte before the quad: every quad restarts its lookup chain from that byte's row
#pragma unroll
					for(int h = 0; h < 4; ++h) {
						const uint4 v = lds128(my_sa + 16u * h);
						const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
						uint32_t packed = 0;
#pragma unroll
						for(int k = 0; k < 4; ++k) {
							const int q = 4 * h + k;
							if(ORDER && q) row = row_of(lastc);
							uint32_t e[4];
#pragma unroll
							for(int j = 0; j < 4; ++j) {
								const uint32_t c = __byte_perm(w4[k], 0, 0x4440 + j);
								uint32_t ent = 0;
								if(!checked || 4 * q + j < live) {
									ent = ORDER ? lds32(row + c * 4) : lds32(row + c * 128);
									if(ORDER) row = table_sa + __byte_perm(ent, 0, 0x4442) * kCtxPitch;
									floor = min(floor, ent);
									ceil = max(ceil, ent);
								}
								e[j] = ent;
							}
							const uint32_t l1 = e[1] >> 27, l3 = e[3] >> 27;
							const uint32_t lp0 = (e[0] >> 27) + l1, lp1 = (e[2] >> 27) + l3;
							const uint32_t p0 = ((e[0] & 0xffffu) << l1) | (e[1] & 0xffffu);
							const uint32_t p1 = ((e[2] & 0xffffu) << l3) | (e[3] & 0xffffu);
							uint32_t lo;
							asm("shl.b32 %0, %1, %2;" : "=r"(lo) : "r"(p0), "r"(lp1));   // lp1 may be 32
							q_lo[q] = lo | p1;
							packed |= (lp0 + lp1) << (8 * k);
							my_bits += lp0 + lp1;
							lastc = w4[k] >> 24;
						}
						qlen4[h] = packed;
					}
				};
				if(live == 64) quads64(false);
				else quads64(true);
				// Rare: a quad longer than 32 bits (its low word alone is not the quad), or a codeword longer than 16 bits somewhere
				// in this thread (length marker 31; only a quad whose merged length reaches 31 can hold one): those quads take
				// their lengths from the wide table and are packed symbol by symbol below.
				uint32_t longq = 0;
#pragma unroll
				for(int h = 0; h < 4; ++h) longq |= (qlen4[h] + 0x5f5f5f5fu) & 0x80808080u;   // some quad of 33 bits or more (a length is at most 4 x 31)
				if((ceil >> 27) == 31u || longq) {
#pragma unroll
					for(int q = 0; q < 16; ++q) {
						const uint32_t len = (qlen4[q >> 2] >> (8 * (q & 3))) & 255u;
						if(len >= 31u) {
							uint32_t wq;   // the quad's bytes are still in the input buffer (a volatile load: lds32 is a pure function of its address to the compiler, right for the table, wrong for a buffer that is refilled every tile)
							asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wq) : "r"(my_sa + 4u * q) : "memory");
							uint32_t p = prev, sum = 0;
							if(q) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p) : "r"(my_sa + 4u * q - 1u) : "memory");
							bool big = false;
#pragma unroll
							for(int i = 0; i < 4; ++i) {
								if(4 * q + i < live) {
									const uint32_t c = __byte_perm(wq, 0, 0x4440 + i);
									const uint32_t l = uint32_t(__ldg(A.wide + ((ORDER ? p : 0u) << 8) + c) >> 56);
									big |= l > 16u;
									sum += l;
									p = c;
								}
							}
							if(big || len > 32u) {
								esc_mask |= 1u << q;
								my_bits += sum - len;
							}
						}
					}
				}
				if(floor < (1u << 27)) {   // rare: some symbol has no codeword; count them exactly (context rows again)
					uint32_t r2 = table_sa + lane * 4;
					if(ORDER) r2 = row_of(prev);
#pragma unroll 1
					for(int i = 0; i < live; ++i) {
						const uint32_t c = A.in[my + i];
						const uint32_t ent = ORDER ? lds32(r2 + c * 4) : lds32(r2 + c * 128);
						if(ORDER) r2 = table_sa + __byte_perm(ent, 0, 0x4442) * kCtxPitch;
						dropped += (ent >> 27) == 0 ? 1u : 0u;
					}
				}
			} else {
			uint32_t w[NWORDS];
#pragma unroll
			for(int k = 0; k < NWORDS; ++k) w[k] = 0;
			int live = 0;
			const bool from_smem = TMA && tile < full_tiles;
			if(from_smem) {   // the tile came by bulk copy: wait for it, then 2 x 16 bytes per thread from shared memory
				mbar_wait(mbar_sa, in_parity);
				in_parity ^= 1u;
				live = SPT;
#pragma unroll
				for(int k = 0; k < SPT / 16; ++k) {
					const uint4 v = lds128(inbuf_sa + kEncInbufLead + tid * SPT + 16 * k);
					w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
				}
			} else if(my < A.n) {
				live = A.n - my >= uint64_t(SPT) ? SPT : int(A.n - my);
				if(ALIGNED && live == SPT) {
#pragma unroll
					for(int k = 0; k < SPT / 16; ++k) {
						const uint4 v = ld_stream_128(A.in + my + 16 * k);
						w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
					}
				} else {
#pragma unroll
					for(int i = 0; i < SPT; ++i)
						if(i < live) w[i >> 2] |= uint32_t(A.in[my + i]) << (8 * (i & 3));
				}
			}
			uint32_t prev = __shfl_up_sync(0xffffffffu, w[NWORDS - 1] >> 24, 1);
			if(lane == 0 && my < A.n) {
				if(my == 0) prev = first_prev;
				else if(TMA && tile < full_tiles) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(prev) : "r"(inbuf_sa + kEncInbufLead + tid * SPT - 1u) : "memory");
				else prev = uint32_t(A.in[my - 1]);
			}
			if constexpr(FMT == FMT_CTX) {
				// One shared-memory lookup per symbol; the entry names the next context's row, so inside a quad the
				// dependent chain is lookup -> byte select -> multiply-add -> lookup. Every quad restarts the chain from
				// the byte before it (its row comes from the null row), so the eight quads of a thread are independent.
				// Codewords are merged on the fly: pairs (<= 32 bits), then quads kept as (hi:lo, length) — a quad
				// longer than 32 bits goes out in two steps.
				const uint32_t rank_sa = uint32_t(__cvta_generic_to_shared(s_rank));
				auto row_of = [&](uint32_t byte) -> uint32_t {
					uint32_t r;
					asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(rank_sa + byte));
					return table_sa + r * kCtxPitch;
				};
				uint32_t row = table_sa + lane * 4;
				if(ORDER) row = row_of(prev);
				uint32_t floor = 0xffffffffu, ceil = 0;
				auto lookup = [&](int i, bool checked) -> uint32_t {
					const uint32_t c = __byte_perm(w[i >> 2], 0, 0x4440 + (i & 3));
					uint32_t ent = 0;
					if(!checked || i < live) {
						ent = ORDER ? lds32(row + c * 4) : lds32(row + c * 128);   // order 0: `row` is the lane's column of the replicated row
						if(ORDER) row = table_sa + __byte_perm(ent, 0, 0x4442) * kCtxPitch;
						floor = min(floor, ent);
						ceil = max(ceil, ent);
					}
					return ent;
				};
				auto quads = [&](bool checked) {
#pragma unroll
					for(int q = 0; q < SPT / 4; ++q) {
						if(ORDER && q) row = row_of(__byte_perm(w[q - 1], 0, 0x4443));
						const uint32_t e0 = lookup(4 * q, checked), e1 = lookup(4 * q + 1, checked), e2 = lookup(4 * q + 2, checked), e3 = lookup(4 * q + 3, checked);
						const uint32_t l1 = e1 >> 27, l3 = e3 >> 27;
						const uint32_t lp0 = (e0 >> 27) + l1, lp1 = (e2 >> 27) + l3;
						const uint32_t p0 = ((e0 & 0xffffu) << l1) | (e1 & 0xffffu);
						const uint32_t p1 = ((e2 & 0xffffu) << l3) | (e3 & 0xffffu);
						uint32_t lo;
						asm("shl.b32 %0, %1, %2;" : "=r"(lo) : "r"(p0), "r"(lp1));   // lp1 may be 32
						q_lo[q] = lo | p1;
						q_hi[q] = __funnelshift_lc(p0, 0u, lp1);
						q_len[q] = lp0 + lp1;
						my_bits += lp0 + lp1;
					}
				};
				if(live == SPT) quads(false);
				else quads(true);
				if((ceil >> 27) == 31u) {   // rare: a codeword longer than 16 bits (length marker 31) somewhere in this thread
					// Only a quad whose merged length reaches 31 can hold the marker; those quads take their lengths from
					// the wide table and are packed symbol by symbol below. The other quads keep their merged codewords.
#pragma unroll
					for(int q = 0; q < SPT / 4; ++q) {
						if(q_len[q] >= 31u) {
							uint32_t p = q ? __byte_perm(w[q ? q - 1 : 0], 0, 0x4443) : prev, sum = 0;
							bool big = false;
#pragma unroll
							for(int i = 0; i < 4; ++i) {
								if(4 * q + i < live) {
									const uint32_t c = __byte_perm(w[q], 0, 0x4440 + i);
									const uint32_t len = uint32_t(__ldg(A.wide + ((ORDER ? p : 0u) << 8) + c) >> 56);
									big |= len > 16u;
									sum += len;
									p = c;
								}
							}
							if(big) {
								esc_mask |= 1u << q;
								my_bits += sum - q_len[q];
							}
						}
					}
				}
				if(floor < (1u << 27)) {   // rare: some symbol has no codeword; count them exactly (context rows again)
					uint32_t r2 = table_sa + lane * 4;
					if(ORDER) r2 = row_of(prev);
#pragma unroll 1
					for(int i = 0; i < live; ++i) {
						const uint32_t c = A.in[my + i];   // re-read: indexing w[] dynamically would push it to local memory
						const uint32_t ent = ORDER ? lds32(r2 + c * 4) : lds32(r2 + c * 128);
						if(ORDER) r2 = table_sa + __byte_perm(ent, 0, 0x4442) * kCtxPitch;
						dropped += (ent >> 27) == 0 ? 1u : 0u;
					}
				}
			} else if constexpr(FMT != FMT_WIDE) {
				// entries: length in [31:27], right-aligned code in [26:0]
				uint32_t floor = 0xffffffffu;
				// byte address of the current context's row (box formats keep the border row at index R)
				const uint32_t pitch4 = pitch * 4;
				uint32_t row = (FMT == FMT_BOX_SMEM ? table_sa : 0u) + (A.order ? min(prev - lo, R) * pitch4 : 0u);
				const bool whole = live == SPT;
#pragma unroll
				for(int i = 0; i < SPT; ++i) {
					const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 255u;
					uint32_t ent = 0;
					if(whole || i < live) {
						if(A.order) {
							const uint32_t uc = min(c - lo, R);   // bytes outside the box land on the zero border
							ent = FMT == FMT_BOX_SMEM ? lds32(row + uc * 4) : __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(A.box) + row) + uc);
							row = (FMT == FMT_BOX_SMEM ? table_sa : 0u) + uc * pitch4;
						} else {
							ent = FMT == FMT_BOX_SMEM ? lds32(row + c * 4) : __ldg(A.box + c);
						}
						floor = min(floor, ent);
					}
					e32[i] = ent;
					my_bits += ent >> 27;
				}
				if(floor == 0) {   // rare: count the symbols without a codeword exactly
#pragma unroll
					for(int i = 0; i < SPT; ++i) dropped += (i < live && e32[i] == 0) ? 1u : 0u;
				}
			} else {
				// u64 entries (codewords up to 56 bits)
#pragma unroll
				for(int i = 0; i < SPT; ++i) {
					const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 255u;
					unsigned long long ent = 0;
					if(i < live) {
						ent = __ldg(A.wide + ((A.order ? prev : 0u) << 8) + c);
						if(ent == 0) ++dropped;
					}
					e64[i] = ent;
					my_bits += uint32_t(ent >> 56);
					prev = c;
				}
			}
			}   // SPT != 64
			// block exclusive scan; the tile's bit count is published right away
			pos = block_exclusive_scan(my_bits, warp_sums, tile_bits);
			if constexpr(SPT == 64) {
				overflow = tile_bits > (stage_words - 7u) * 32u;
				if(overflow && tid == 0) *A.fail = 1u;   // this launch's output is void: the SPT = 32 launch behind it encodes the stream again
			}
			if(tid == 0) {
				st_relaxed(A.agg + tile, kAgg | tile_bits);
				s_pub_tile[it & 1] = tile;
				s_pub_bits[it & 1] = tile_bits;
			}
			bar_arrive(2 + (it & 1));   // the scanner takes it from here
			// the next ticket is visible since the scan's barrier: pull that tile's input into L2 while this one packs
			const uint32_t next_tile = s_tile[(it + 1) & 1];
			if((!TMA || next_tile >= full_tiles) && (tid & 3) == 0 && next_tile < A.n_tiles) {
				const uint64_t off = uint64_t(next_tile) * (kEncThreads * SPT) + uint64_t(tid) * SPT;
				if(off < A.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.in + off));
			}
		}

		if(!valid) {   // no tile left: tell the scanner, then write out the pending one
			if(tid == 0) s_pub_tile[it & 1] = 0xffffffffu;
			bar_arrive(2 + (it & 1));
		}

		// ================= phase B: write out the pending tile (packed in iteration it - 1) =================
		if(pending) {
			bar_wait(4 + ((it - 1) & 1));   // normally passed at once: the scanner had a whole iteration
			const unsigned long long prefix_bits = s_prefix_bits[(it - 1) & 1];
			// funnel-shift copy-out
			const unsigned long long g0 = bit0 + prefix_bits;     // global bit index of the tile's first bit
			const unsigned long long g1 = g0 + p_bits;
			const uint32_t s = uint32_t(g0 & 31);
			const unsigned long long w0 = g0 >> 5;
			unsigned long long w1 = g1 >> 5;                          // words [w0, w1) are completed by this tile
			if(p_tile == A.n_tiles - 1 && (g1 & 31)) ++w1;            // the stream's last partial word, zero padded
			const uint32_t nw = uint32_t(w1 - w0);
			if(w1 > A.out_capacity_words) {
				if(tid == 0) A.result[2] = 1;                         // capacity error; this tile writes nothing
			} else {
				// a word shared with the predecessor (s != 0: its last s bits open word w0) is the scanner's
				uint32_t j = tid + (s ? 1u : 0u);
				uint32_t* dp = A.out_words + w0 + j;
				const uint32_t* sp = stage + j;
#pragma unroll 1
				for(; j < nw; j += kEncThreads, dp += kEncThreads, sp += kEncThreads)
					*dp = __byte_perm(__funnelshift_r(sp[0], sp[-1], s), 0, 0x0123);   // sp[-1] of word 0 is the zero word in front
			}
			bar_workers();
			const uint32_t used4 = ((p_bits + 31) / 32 + 1 + 3) / 4;   // 16 bytes per store; the staging area has the slack
#pragma unroll 1   // one trip for text (a tile stages ~400 such stores): unrolled, the trip-count arithmetic cost more than the stores
			for(uint32_t j = tid; j < used4; j += kEncThreads) reinterpret_cast<uint4*>(stage)[j] = make_uint4(0, 0, 0, 0);
			bar_workers();   // the staging area is clean before the next tile is packed into it
			pending = false;
		}

		// ================= phase C: pack tile `tile` into the staging area =================
		if(valid) {
			if constexpr(SPT == 64) {
				if(!overflow) {
					Packer32 pk;
					pk.start(stage_sa, pos);
#pragma unroll
					for(int q = 0; q < 16; ++q) {
						if(esc_mask & (1u << q)) {   // rare: a quad longer than 32 bits or with a codeword of 17..29 bits: one unit per symbol, from the wide table
							const uint64_t at = my + 4 * q;
							uint32_t p = at == 0 ? first_prev : uint32_t(A.in[at - 1]);
#pragma unroll 1
							for(int i = 0; i < 4 && at + i < A.n; ++i) {
								const uint32_t c = A.in[at + i];
								const unsigned long long ent = __ldg(A.wide + ((ORDER ? p : 0u) << 8) + c);
								pk.put(uint32_t(ent), uint32_t(ent >> 56));
								p = c;
							}
						} else {
							pk.put(q_lo[q], (qlen4[q >> 2] >> (8 * (q & 3))) & 255u);
						}
					}
					pk.finish();
				}
			} else if constexpr(FMT == FMT_CTX) {
				Packer32 pk;
				pk.start(stage_sa, pos);
#pragma unroll
				for(int q = 0; q < SPT / 4; ++q) {
					if(esc_mask & (1u << q)) {   // rare: the quad holds a codeword of 17..28 bits: one unit per symbol, from the wide table
						const uint64_t at = my + 4 * q;
						uint32_t p = at == 0 ? first_prev : uint32_t(A.in[at - 1]);
#pragma unroll 1
						for(int i = 0; i < 4 && at + i < A.n; ++i) {
							const uint32_t c = A.in[at + i];
							const unsigned long long ent = __ldg(A.wide + ((ORDER ? p : 0u) << 8) + c);
							pk.put(uint32_t(ent), uint32_t(ent >> 56));
							p = c;
						}
					} else if(q_len[q] <= 32) {
						pk.put(q_lo[q], q_len[q]);
					} else {   // a quad of long codewords: the top q_len - 32 bits, then the low word
						pk.put(q_hi[q], q_len[q] - 32);
						pk.put(q_lo[q], 32);
					}
				}
				pk.finish();
			}
			Packer pk;
			pk.start(stage_sa, pos);
			if constexpr(FMT == FMT_CTX) {
			} else if constexpr(FMT != FMT_WIDE) {
				// merge pairs -> quads, funnel through the window
#pragma unroll
				for(int q = 0; q < SPT / 4; ++q) {
					const uint32_t l0 = e32[4 * q] >> 27, l1 = e32[4 * q + 1] >> 27, l2 = e32[4 * q + 2] >> 27, l3 = e32[4 * q + 3] >> 27;
					const unsigned long long p0 = (uint64_t(e32[4 * q] & 0x07ffffffu) << l1) | (e32[4 * q + 1] & 0x07ffffffu);
					const unsigned long long p1 = (uint64_t(e32[4 * q + 2] & 0x07ffffffu) << l3) | (e32[4 * q + 3] & 0x07ffffffu);
					const uint32_t lp0 = l0 + l1, lp1 = l2 + l3;
					if(lp0 + lp1 <= 64) {
						pk.put((p0 << lp1) | p1, lp0 + lp1);
					} else {   // four long codewords in a row: two <= 54-bit units
						pk.put(p0, lp0);
						pk.put(p1, lp1);
					}
				}
			} else {
#pragma unroll
				for(int i = 0; i < SPT; ++i) pk.put(e64[i] & 0x00ffffffffffffffull, uint32_t(e64[i] >> 56));
			}
			if constexpr(FMT != FMT_CTX) pk.finish();
			bar_workers();   // staged bits visible
			if(tid == 0) {     // the tail goes out now: successors need it only when they write their first word
				const uint32_t tcount = tile_bits < 31 ? tile_bits : 31;
				st_relaxed32(A.tail + tile, kTailValid | (overflow ? 0u : stage_bits(stage, tile_bits - tcount, tcount)));
				s_head[it & 1] = stage[0];   // for the scanner: the bits that share an output word with the predecessor
			}
			pending = true;
			p_tile = tile;
			p_bits = overflow ? 0u : tile_bits;   // (a tile that did not fit was not packed: nothing to copy out)
		}
	}
	dropped = uint32_t(warp_sum(dropped));
	if(lane == 0 && dropped) atomicAdd(A.result + 1, (unsigned long long) dropped);
}

// ---------------------------------------------------------------------------------------------------------
// K2w: the context-row encoder with WARP-PRIVATE tiles (round 2; opt-in with the tunable enc_warp = 1).
// One CTA of 32 warps per SM; after the table is staged there is no CTA-wide barrier, no per-CTA scanner warp and no
// block scan: every warp pulls its own tiles of 32 x 32 input bytes from the atomic ticket and takes each of them through
// the whole pipeline by itself, so the 32 warps of an SM are independent instruction streams:
//   input     one bulk copy (TMA) per tile into the warp's own shared-memory buffer, issued by lane 0 one tile ahead,
//             completion through the warp's own mbarrier; or 2 x 128-bit loads per lane (unaligned input, ragged tile)
//   lookup    as in encode_kernel: context rows, eight independent quads per lane, merged on the fly
//   scan      warp shuffles; the tile's bit count goes out at once: a relaxed store of the tile's aggregate word and one
//             RED into the accumulator of its GROUP of 64 consecutive tiles (tiles counted << 48 | bits)
//   pack      into the warp's own zeroed staging area (shared-memory OR of completed words), then the tile's last 31 bits
//             are published (the successor needs them for the output word the two tiles share)
//   prefix    ONE scan warp in the whole grid (warp 31 of the CTA that wins an election, so it is certainly running) turns
//             complete group accumulators into inclusive prefixes, 32 groups per step, in order
//   write-out one iteration later (software pipeline as in encode_kernel): bits before the tile = the scan warp's prefix of
//             the group before (one word) + the aggregates of the lower tiles of its own group (one poll); funnel shift by
//             the global bit phase, byte swap, coalesced 32-bit stores; the word shared with the predecessor is completed
//             with the predecessor's tail bits, so every output word still has one writer; the staging words are
//             re-zeroed in the same pass
// A tile only ever waits for tiles with smaller tickets, which are running: no deadlock whatever the residency.
//
// What it measured (1 GiB Markov text, same box, alternating): 1.145 ms against 1.203 ms for encode_kernel — and
// 1.22 / 1.10 ms with -h, 0.36 / 0.33 and 0.49 / 0.34 ms on the 256 MiB Fibonacci streams, so it stays opt-in. On the way:
//   * a decoupled look-back per warp tile: 5.4 ms — with ~4700 tiles in flight the nearest inclusive prefix is thousands of
//     tiles back (148 polls of 32); with groups and a look-back over group descriptors 1.38 ms (5.6 polls per tile, and
//     every warp polls the same few L2 lines); walkers that publish what they found: 14 ms (stores into polled lines);
//   * a ticket worth two tiles: 377 ms — the second tile stays reserved but unstarted for a tile time, its bit count comes
//     as late as its successors need it, every wait delays the next publication and the delays feed each other;
//   * one coordinator warp per CTA, chunks of 31 tiles per ticket, mbarrier hand-offs (chunk / bit counts / base) in
//     shared memory, no global protocol traffic from the workers: 1.18 ms (the workers wait for the next chunk);
//   * two staging areas per warp, write-out two iterations later (no bulk-copied input then: shared memory): the waits
//     for the prefix disappear, 1.20-1.37 ms.
// The instruction count per tile is no lower than encode_kernel's (about 1000 against 897 warp instructions per 1024
// symbols: the per-tile bookkeeping is paid per warp instead of per CTA), and what is gained on barriers is lost on
// L2 round trips (ticket, prefix, tail, group poll): this path is bound by its instruction count, not by its barriers.
// ---------------------------------------------------------------------------------------------------------
constexpr int kEncWarpThreads = 1024;
constexpr uint32_t kWarpTileBytes = 32 * 32;
constexpr uint32_t kEncGroupTiles = 64;                    // tiles per look-back group
constexpr uint32_t kGroupAccStride = 16;                   // u64 words between two groups' accumulators: a 128-byte line each
constexpr unsigned long long kGroupBitsMask = (1ull << 48) - 1;   // group accumulator: tiles counted << 48 | bits
constexpr unsigned long long kTail64Valid = 1ull << 63;   // tail word: valid | min(bits, 31) << 32 | the tile's last bits

__host__ __device__ __forceinline__ uint32_t warp_stage_words(uint32_t longest) {   // staged words of one warp tile, whatever the input
	return ((kWarpTileBytes * (longest ? longest : 1u) + 31u) / 32u + 3u) & ~3u;
}
// dynamic shared memory: table, 32 x (4 zero words + stage_words + 4 words of slack), 32 input buffers (TMA)
__host__ __device__ __forceinline__ uint32_t warp_kernel_smem(uint32_t ctx_rows, uint32_t stage_words, bool tma) {
	return ((ctx_rows * kCtxPitch + 15u) & ~15u) + (kEncWarpThreads / 32) * (stage_words + 8u) * 4u + (tma ? (kEncWarpThreads / 32) * (kWarpTileBytes + kEncInbufLead) : 0u);
}

// The last min(bits before this tile, 31) bits of the stream before `tile` (one lane).
__device__ __forceinline__ uint32_t predecessor_tail64(uint32_t tile, const unsigned long long* tails) {
	if(tile == 0) return 0;
	unsigned long long tv;
	do { tv = ld_relaxed(tails + tile - 1); } while(!(tv & kTail64Valid));
	uint32_t t_tail = uint32_t(tv) & uint32_t(kLow31), t_bits = uint32_t(tv >> 32) & 63u;
	// Rare (a predecessor produced fewer than 31 bits: dropped symbols): keep prepending older tiles.
	for(long long j = (long long) tile - 2; j >= 0 && t_bits < 31; --j) {
		do { tv = ld_relaxed(tails + j); } while(!(tv & kTail64Valid));
		const uint32_t b = uint32_t(tv >> 32) & 63u;
		t_tail = uint32_t(((uint64_t(uint32_t(tv) & uint32_t(kLow31)) << t_bits) | t_tail) & kLow31);
		t_bits = b + t_bits > 31 ? 31u : b + t_bits;
	}
	return t_tail;
}

template <int ORDER, bool TMA>
__global__ void __launch_bounds__(kEncWarpThreads, 1) encode_warp_kernel(const EncArgs A) {
	constexpr int SPT = 32, NWORDS = SPT / 4;
	constexpr uint32_t NW = kEncWarpThreads / 32;
	__shared__ uint8_t s_rank[256];   // row of every byte value as a context
	__shared__ __align__(8) unsigned long long s_mbar[NW];
	extern __shared__ uint32_t smem[];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t bit0 = A.bit_base_dev ? uint32_t(__ldg(A.bit_base_dev) & 7ull) : A.bit0;
	const uint32_t first_prev = A.prev0_dev ? uint32_t(__ldg(A.prev0_dev) & 255ull) : A.prev0;
	uint32_t ctx_rows = A.ctx_rows, stage_words = A.stage_words;
	if(A.meta) {   // device-built tables: check that this launch's shared memory fits them
		ctx_rows = __ldg(A.meta);
		const uint32_t status = __ldg(A.meta + 1), longest = __ldg(A.meta + 2);
		stage_words = warp_stage_words(longest);
		if(status != 0 || ctx_rows > uint32_t(kEncCtxMaxRows) || longest > uint32_t(kEncCtxMaxBits) || warp_kernel_smem(ctx_rows, stage_words, TMA) > A.smem_bytes) {
			if(blockIdx.x == 0 && tid == 0) A.result[3] = status ? (unsigned long long) (long long) (int) status : 1ull;   // the caller takes the host-built path
			return;
		}
	}
	for(uint32_t i = tid; i < ctx_rows * 256u; i += kEncWarpThreads) smem[(i >> 8) * (kCtxPitch / 4) + (i & 255u)] = __ldg(A.ctx + i);
	if(tid < 256) s_rank[tid] = uint8_t(__ldg(A.ctx + (ctx_rows - 1) * 256u + tid) >> 16);
	const uint32_t table_words = ((ctx_rows * kCtxPitch + 15u) & ~15u) / 4u;
	const uint32_t per_warp = stage_words + 8u;
	for(uint32_t i = tid; i < per_warp * NW; i += kEncWarpThreads) smem[table_words + i] = 0;
	const uint32_t table_sa = uint32_t(__cvta_generic_to_shared(smem));
	const uint32_t rank_sa = uint32_t(__cvta_generic_to_shared(s_rank));
	uint32_t* stage = smem + table_words + warp * per_warp + 4;   // stage[-1] reads as "no bits"
	const uint32_t stage_sa = uint32_t(__cvta_generic_to_shared(stage));
	const uint32_t inbuf_sa = table_sa + (table_words + per_warp * NW) * 4u + warp * (kWarpTileBytes + kEncInbufLead);
	const uint32_t mbar_sa = uint32_t(__cvta_generic_to_shared(&s_mbar[warp]));
	const uint32_t full_tiles = A.full_tiles;
	if(TMA && lane == 0) {
		mbar_init(mbar_sa, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();   // the only CTA-wide barrier

	auto issue_tile = [&](uint32_t t) {   // lane 0: the tile's bytes and the 16 before them (its first context)
		if(t >= full_tiles) return;
		const uint64_t at = uint64_t(t) * kWarpTileBytes;
		if(t) bulk_load_tile(inbuf_sa, A.in + at - kEncInbufLead, kWarpTileBytes + kEncInbufLead, mbar_sa);
		else bulk_load_tile(inbuf_sa + kEncInbufLead, A.in, kWarpTileBytes, mbar_sa);
	};
	auto row_of = [&](uint32_t byte) -> uint32_t {
		uint32_t r;
		asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(rank_sa + byte));
		return table_sa + r * kCtxPitch;
	};

	// ---- the scan warp: ONE warp of the whole grid (warp 31 of the CTA that wins the election, so it is certainly running)
	// turns the groups' complete accumulators into inclusive prefixes, 32 groups per step, in order. The tiles then read
	// ONE word each. (Tiles that walk back over the group descriptors themselves — a decoupled look-back per warp — were
	// 15 % slower at best: with ~4700 tiles in flight everybody polls the same few L2 lines, and the walk is 5 polls long.)
	if(warp == NW - 1) {
		uint32_t won = 0;
		if(lane == 0) won = atomicCAS(A.ticket + 1, 0u, 1u) == 0u ? 1u : 0u;
		if(__shfl_sync(0xffffffffu, won, 0)) {
			const uint32_t n_groups = (A.n_tiles + kEncGroupTiles - 1) / kEncGroupTiles;
			unsigned long long running = 0;
			for(uint32_t g0 = 0; g0 < n_groups;) {
				const uint32_t g = g0 + lane;
				const uint32_t want = g < n_groups ? (A.n_tiles - g * kEncGroupTiles < kEncGroupTiles ? A.n_tiles - g * kEncGroupTiles : kEncGroupTiles) : 0xffffffffu;
				const unsigned long long acc = g < n_groups ? ld_relaxed(A.grp_acc + size_t(g) * kGroupAccStride) : 0ull;
				const unsigned open = __ballot_sync(0xffffffffu, uint32_t(acc >> 48) != want);   // lanes beyond the last group always count as open
				const uint32_t done = open ? uint32_t(__ffs(open) - 1) : 32u;   // complete groups in front
				if(done == 0) { __nanosleep(100); continue; }   // poll again
				unsigned long long incl = acc & kGroupBitsMask;
#pragma unroll
				for(int d = 1; d < 32; d <<= 1) {
					const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
					if(lane >= uint32_t(d)) incl += t;
				}
				if(lane < done) st_relaxed(A.grp_inc + g, kTail64Valid | (running + incl));
				running += __shfl_sync(0xffffffffu, incl, done - 1);
				g0 += done;
			}
			return;
		}
	}

	uint32_t dropped = 0, in_parity = 0;
	uint32_t tile = 0;
	if(lane == 0) {
		tile = atomicAdd(A.ticket, 1u);
		if(TMA && tile < A.n_tiles) issue_tile(tile);
	}
	tile = __shfl_sync(0xffffffffu, tile, 0);
	// Software pipeline, as in encode_kernel: a tile is written out one iteration after it was packed — after the NEXT
	// tile has been loaded, looked up, counted and published. By then every tile with a smaller ticket has long published
	// its bit count (a tile publishes well within one tile time of its ticket, the look-back runs two tile times after it),
	// so the look-back never waits; and the next tile's bit count is out as early as it can be.
	bool pending = false;   // a packed tile is waiting in the staging area
	uint32_t p_tile = 0, p_bits = 0;
	for(;;) {
		const bool valid = tile < A.n_tiles;
		if(!valid && !pending) break;
		uint32_t next_ticket = tile;   // lane 0 only; the atomic's answer is first looked at after the lookups
		// (one tile per ticket: a ticket worth two tiles keeps the second one reserved but unstarted for a whole tile time, its
		// bit count comes as late as its successors' look-back, every look-back waits a little, the delays feed each other
		// and the kernel all but serialises: measured 377 ms)
		if(valid && lane == 0) next_ticket = atomicAdd(A.ticket, 1u);
		const uint64_t my = uint64_t(tile) * kWarpTileBytes + lane * SPT;
		uint32_t q_hi[SPT / 4], q_lo[SPT / 4], q_len[SPT / 4];
		uint32_t esc_mask = 0, pos = 0, tile_bits = 0;
		if(valid) {
		// ================= phase A: load, look up, count, publish (tile `tile`) =================
		// ---- input ----
		uint32_t w[NWORDS];
#pragma unroll
		for(int k = 0; k < NWORDS; ++k) w[k] = 0;
		int live = 0;
		uint32_t prev = 0;
		const bool from_smem = TMA && tile < full_tiles;
		if(from_smem) {
			mbar_wait(mbar_sa, in_parity);
			in_parity ^= 1u;
			live = SPT;
#pragma unroll
			for(int k = 0; k < SPT / 16; ++k) {
				const uint4 v = lds128(inbuf_sa + kEncInbufLead + lane * SPT + 16 * k);
				w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
			}
			if(lane == 0 && tile) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(prev) : "r"(inbuf_sa + kEncInbufLead - 1u) : "memory");
		} else if(my < A.n) {
			live = A.n - my >= uint64_t(SPT) ? SPT : int(A.n - my);
			if(live == SPT && ((reinterpret_cast<uint64_t>(A.in) & 15) == 0)) {
#pragma unroll
				for(int k = 0; k < SPT / 16; ++k) {
					const uint4 v = ld_stream_128(A.in + my + 16 * k);
					w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
				}
			} else {
#pragma unroll
				for(int i = 0; i < SPT; ++i)
					if(i < live) w[i >> 2] |= uint32_t(A.in[my + i]) << (8 * (i & 3));
			}
			if(lane == 0 && my) prev = uint32_t(A.in[my - 1]);
		}
		{
			const uint32_t up = __shfl_up_sync(0xffffffffu, w[NWORDS - 1] >> 24, 1);
			if(lane) prev = up;
			else if(my == 0) prev = first_prev;
		}

		// ---- lookup + merge: quads as (hi:lo, length); see encode_kernel ----
		uint32_t my_bits = 0;
		{
			uint32_t row = table_sa;
			if(ORDER) row = row_of(prev);
			uint32_t floor = 0xffffffffu, ceil = 0;
			auto lookup = [&](int i, bool checked) -> uint32_t {
				const uint32_t c = __byte_perm(w[i >> 2], 0, 0x4440 + (i & 3));
				uint32_t ent = 0;
				if(!checked || i < live) {
					ent = lds32(row + c * 4);
					if(ORDER) row = table_sa + __byte_perm(ent, 0, 0x4442) * kCtxPitch;
					floor = min(floor, ent);
					ceil = max(ceil, ent);
				}
				return ent;
			};
			auto quads = [&](bool checked) {
#pragma unroll
				for(int q = 0; q < SPT / 4; ++q) {
					if(ORDER && q) row = row_of(__byte_perm(w[q - 1], 0, 0x4443));
					const uint32_t e0 = lookup(4 * q, checked), e1 = lookup(4 * q + 1, checked), e2 = lookup(4 * q + 2, checked), e3 = lookup(4 * q + 3, checked);
					const uint32_t l1 = e1 >> 27, l3 = e3 >> 27;
					const uint32_t lp0 = (e0 >> 27) + l1, lp1 = (e2 >> 27) + l3;
					const uint32_t p0 = ((e0 & 0xffffu) << l1) | (e1 & 0xffffu);
					const uint32_t p1 = ((e2 & 0xffffu) << l3) | (e3 & 0xffffu);
					uint32_t lo;
					asm("shl.b32 %0, %1, %2;" : "=r"(lo) : "r"(p0), "r"(lp1));   // lp1 may be 32
					q_lo[q] = lo | p1;
					q_hi[q] = __funnelshift_lc(p0, 0u, lp1);
					q_len[q] = lp0 + lp1;
					my_bits += lp0 + lp1;
				}
			};
			if(live == SPT) quads(false);
			else quads(true);
			if((ceil >> 27) == 31u) {   // rare: a codeword longer than 16 bits (length marker 31) somewhere in this lane
#pragma unroll
				for(int q = 0; q < SPT / 4; ++q) {
					if(q_len[q] >= 31u) {
						uint32_t p = q ? __byte_perm(w[q ? q - 1 : 0], 0, 0x4443) : prev, sum = 0;
						bool big = false;
#pragma unroll
						for(int i = 0; i < 4; ++i) {
							if(4 * q + i < live) {
								const uint32_t c = __byte_perm(w[q], 0, 0x4440 + i);
								const uint32_t len = uint32_t(__ldg(A.wide + ((ORDER ? p : 0u) << 8) + c) >> 56);
								big |= len > 16u;
								sum += len;
								p = c;
							}
						}
						if(big) {
							esc_mask |= 1u << q;
							my_bits += sum - q_len[q];
						}
					}
				}
			}
			if(floor < (1u << 27)) {   // rare: some symbol has no codeword; count them exactly (context rows again)
				uint32_t r2 = table_sa;
				if(ORDER) r2 = row_of(prev);
#pragma unroll 1
				for(int i = 0; i < live; ++i) {
					const uint32_t c = A.in[my + i];
					const uint32_t ent = lds32(r2 + c * 4);
					if(ORDER) r2 = table_sa + __byte_perm(ent, 0, 0x4442) * kCtxPitch;
					dropped += (ent >> 27) == 0 ? 1u : 0u;
				}
			}
		}
		// ---- the next tile's input: every lane has long taken this tile's bytes out of the buffer ----
		if(TMA && lane == 0) issue_tile(next_ticket);
		if(!TMA) {   // no bulk copy: pull it into L2 while this tile is packed
			const uint32_t nt = __shfl_sync(0xffffffffu, next_ticket, 0);
			const uint64_t off = uint64_t(nt) * kWarpTileBytes + uint64_t(lane) * 128u;
			if(lane < kWarpTileBytes / 128 && nt < A.n_tiles && off < A.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.in + off));
		}
		// ---- warp scan; the tile's bit count is published right away ----
		uint32_t incl = my_bits;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
			if(lane >= uint32_t(d)) incl += t;
		}
		tile_bits = __shfl_sync(0xffffffffu, incl, 31);
		pos = incl - my_bits;
		// Tiles come in groups of kEncGroupTiles consecutive tickets. Thousands of warp tiles are in flight at once, and a
		// look-back over single tiles would have to walk back through all of them to the nearest inclusive prefix (measured:
		// 5.4 ms per GiB); with groups it reads its own group's tiles (one poll) and ~20 group descriptors (one poll).
		// A group's aggregate is complete when its accumulator has counted all its tiles (no return value: a RED).
		if(lane == 0) {
			st_relaxed(A.agg + tile, kAgg | tile_bits);
			atomicAdd(A.grp_acc + size_t(tile / kEncGroupTiles) * kGroupAccStride, (1ull << 48) | tile_bits);   // a RED: nobody waits for it
		}
		}   // phase A
		const uint32_t next_tile = __shfl_sync(0xffffffffu, next_ticket, 0);

		// ================= phase B: write out the pending tile (packed in the previous iteration) =================
		if(pending) {
		const uint32_t group = p_tile / kEncGroupTiles, gi = p_tile % kEncGroupTiles, gbase = group * kEncGroupTiles;
		// ---- look-back, two levels: the lower tiles of this tile's group, then the groups before it ----
		unsigned long long excl;
		{
			for(;;) {   // (a) tiles gbase .. tile - 1 (smaller tickets: running, and they publish before they wait for anyone)
				const unsigned long long v0 = lane < gi ? ld_relaxed(A.agg + gbase + lane) : kAgg;
				const unsigned long long v1 = lane + 32 < gi ? ld_relaxed(A.agg + gbase + 32 + lane) : kAgg;
				if(!__all_sync(0xffffffffu, (v0 >> 62) != 0 && (v1 >> 62) != 0)) { __nanosleep(100); continue; }   // poll again
				excl = warp_sum((v0 & kDescValueMask) + (v1 & kDescValueMask));
				break;
			}
			if(group) {   // (b) the bits before this group: the scan warp's running prefix (one word, normally there long since)
				unsigned long long inc = 0;
				if(lane == 0) while(!((inc = ld_relaxed(A.grp_inc + group - 1)) >> 63)) __nanosleep(200);
				excl += __shfl_sync(0xffffffffu, inc, 0) & ~kTail64Valid;
			}
			if(lane == 0) {
				if(p_tile == A.n_tiles - 1) A.result[0] = excl + p_bits;
			}
		}

		// ---- copy-out: funnel shift by the global bit phase; the staging words are re-zeroed on the way ----
		{
			const unsigned long long g0 = bit0 + excl, g1 = g0 + p_bits;
			const uint32_t s = uint32_t(g0 & 31);
			const unsigned long long w0 = g0 >> 5;
			unsigned long long w1 = g1 >> 5;                               // words [w0, w1) are completed by this tile
			if(p_tile == A.n_tiles - 1 && (g1 & 31)) ++w1;                   // the stream's last partial word, zero padded
			uint32_t nw = uint32_t(w1 - w0);
			if(w1 > A.out_capacity_words) {
				if(lane == 0) A.result[2] = 1;                             // capacity error; this tile writes nothing
				nw = 0;
			}
			const uint32_t used = (p_bits + 31) >> 5;                   // staging words that hold bits (nw <= used + 1)
			// the word shared with the predecessor: its last s bits open word w0
			uint32_t carry = 0;
			if(lane == 0 && s && nw) carry = predecessor_tail64(p_tile, A.inc) & ((1u << s) - 1u);
			uint32_t* dp = A.out_words + w0 + lane;
			for(uint32_t j = lane; j - lane <= used; j += 32, dp += 32) {
				const uint32_t cur = j <= used ? lds32(stage_sa + j * 4) : 0u;   // word `used` is a zero word
				uint32_t before = __shfl_up_sync(0xffffffffu, cur, 1);
				if(lane == 0) before = carry;
				carry = __shfl_sync(0xffffffffu, cur, 31);
				if(j < nw) *dp = __byte_perm(__funnelshift_r(cur, before, s), 0, 0x0123);
				if(j < used) stage[j] = 0;
			}
		}
		__syncwarp();   // the staging area is clean before the next tile is packed into it
		pending = false;
		}   // phase B

		// ================= phase C: pack tile `tile` into the staging area =================
		if(valid) {
		// ---- pack into the warp's staging area ----
		{
			Packer32 pk;
			pk.start(stage_sa, pos);
#pragma unroll
			for(int q = 0; q < SPT / 4; ++q) {
				if(esc_mask & (1u << q)) {   // rare: the quad holds a codeword of 17..29 bits: one unit per symbol, from the wide table
					const uint64_t at = my + 4 * q;
					uint32_t p = at == 0 ? first_prev : uint32_t(A.in[at - 1]);
#pragma unroll 1
					for(int i = 0; i < 4 && at + i < A.n; ++i) {
						const uint32_t c = A.in[at + i];
						const unsigned long long ent = __ldg(A.wide + ((ORDER ? p : 0u) << 8) + c);
						pk.put(uint32_t(ent), uint32_t(ent >> 56));
						p = c;
					}
				} else if(q_len[q] <= 32) {
					pk.put(q_lo[q], q_len[q]);
				} else {   // a quad of long codewords: the top q_len - 32 bits, then the low word
					pk.put(q_hi[q], q_len[q] - 32);
					pk.put(q_lo[q], 32);
				}
			}
			pk.finish();
		}
		__syncwarp();   // staged bits visible to the whole warp
		if(lane == 0) {   // the tail goes out now: the successor needs it when it writes its first word, an iteration from now
			const uint32_t tcount = tile_bits < 31 ? tile_bits : 31;
			st_relaxed(A.inc + tile, kTail64Valid | (uint64_t(tcount) << 32) | stage_bits(stage, tile_bits - tcount, tcount));
		}

		pending = true;
		p_tile = tile;
		p_bits = tile_bits;
		}   // phase C
		tile = next_tile;
	}
	dropped = uint32_t(warp_sum(dropped));
	if(lane == 0 && dropped) atomicAdd(A.result + 1, (unsigned long long) dropped);
}

template <typename K>
int launch_kernel(K kern, const EncArgs& args, size_t smem_bytes, cudaStream_t st, const char* prof_name) {
	MH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
	int per_sm = 0;
	MH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEncCtaThreads, smem_bytes));
	if(per_sm < 1) per_sm = 1;
	uint64_t grid = uint64_t(sm_count()) * per_sm;
	if(grid > args.n_tiles) grid = args.n_tiles;
	{
		ProfScope p(prof_name, st);
		kern<<<unsigned(grid), kEncCtaThreads, smem_bytes, st>>>(args);
	}
	count_launch(1);
	MH_CUDA(cudaGetLastError());
	return MH_OK;
}

template <int SPT, int FMT, int ORDER = 1, bool TMA = false>
int launch_variant(bool aligned, const EncArgs& args, size_t smem_bytes, cudaStream_t st, const char* prof_name = "encode_kernel") {
	auto kern = aligned ? encode_kernel<SPT, FMT, true, ORDER, TMA> : encode_kernel<SPT, FMT, false, ORDER, false>;
	return launch_kernel(kern, args, smem_bytes, st, prof_name);
}

}  // namespace

uint64_t encode_tiles_for(uint64_t n) { return (n + kWarpTileBytes - 1) / kWarpTileBytes + 1; }   // the smallest tiles any variant uses (K2w: 1 KiB)

int launch_encode(const uint8_t* d_in, uint64_t n, uint8_t prev0, const mh_codebook* cb, uint64_t bit_base,
                  uint8_t* d_out, uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws, cudaStream_t st,
                  const unsigned long long* d_bit_base, const unsigned long long* d_prev0) {
	if(!cb || !cb->d_enc || !d_result || (!d_in && n) || (!d_out && n)) return MH_ERR_INVALID_ARG;
	if(reinterpret_cast<uint64_t>(d_out) & 3) return MH_ERR_INVALID_ARG;
	if(!ws || !ws->enc_desc) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(d_result, 0, 4 * sizeof(unsigned long long), st));
	if(n == 0) return MH_OK;

	// table format
	// Device-built tables: the host knows neither the rows nor the longest codeword. The launch gets the shared memory
	// that still leaves two CTAs per SM (kEncCtxSmemLimit) and the tile size for kEncCtxMaxBits-bit codewords; the kernel
	// lays table, staging area and input tile out from the real values (d_meta) and reports through d_result[3] when
	// they do not fit — the caller then takes the host-built path.
	const bool dev = cb->device_built;
	const int maxb = dev ? kEncCtxMaxBits : (cb->max_bits > 0 ? cb->max_bits : 1);
	const int force_fmt = int(tunable(kTunEncFmt));   // experiments / tests: force a table format (0: box in shared memory, 1: box in global, 2: wide)
	int fmt = FMT_WIDE;
	size_t table_bytes = 0;
	if(cb->has_box && force_fmt != FMT_WIDE) {
		table_bytes = size_t(cb->order ? (cb->box_r + 1) * (cb->box_r + 1) : 256) * 4;
		fmt = table_bytes <= size_t(kEncBoxSmemLimit) && force_fmt != FMT_BOX_GLOBAL ? FMT_BOX_SMEM : FMT_BOX_GLOBAL;
	}
	if(fmt != FMT_BOX_SMEM) table_bytes = 0;
	// context rows (live contexts only, <= 16-bit codewords): preferred whenever table + staging fit two CTAs per SM
	if(cb->ctx_rows && force_fmt < 0 && size_t(ctx_table_bytes(cb->ctx_rows, cb->order)) + (size_t(kEncThreads) * maxb + 64) * 4 <= size_t(kEncCtxSmemLimit)) {
		fmt = FMT_CTX;
		table_bytes = size_t(ctx_table_bytes(cb->ctx_rows, cb->order));
	}
	if(dev) fmt = FMT_CTX;
	// Tile size: the staged bits of one tile must fit the staging area whatever the input, so symbols per tile x
	// longest codeword bounds it: 32 symbols per thread while that bound stays within kEncStageMaxWords.
	int spt = 16;
	if(fmt != FMT_WIDE && uint64_t(kEncThreads) * 32 * maxb <= uint64_t(kEncStageMaxWords) * 32) spt = 32;
	const uint32_t stage_words = uint32_t(((uint64_t(kEncThreads) * spt * maxb + 31) / 32 + 7) & ~uint64_t(3));
	const uint64_t tile_bytes = uint64_t(kEncThreads) * spt;
	const uint64_t tiles = (n + tile_bytes - 1) / tile_bytes;
	if(tiles > ws->enc_tiles_cap || tiles > 0x7fffffffull) return MH_ERR_WORKSPACE;
	// descriptors: aggregate words, inclusive words, tail words (the warp-private kernel lays its own out below)
	const bool warp_tiles = fmt == FMT_CTX && tunable(kTunEncWarp) > 0;   // opt-in (enc_warp = 1): see the K2w header for what it measured
	if(!warp_tiles) MH_CUDA(cudaMemsetAsync(ws->enc_desc, 0, tiles * (2 * sizeof(uint64_t) + sizeof(uint32_t)), st));
	MH_CUDA(cudaMemsetAsync(ws->counters, 0, 2 * sizeof(uint32_t), st));   // the tile ticket, the scan warp's election flag

	EncArgs a;
	a.in = d_in; a.n = n; a.prev0 = prev0; a.order = cb->order;
	a.wide = reinterpret_cast<const unsigned long long*>(cb->d_enc);
	a.box = cb->d_box; a.box_lo = cb->box_lo; a.box_r = cb->box_r;
	a.ctx = cb->d_ctx; a.ctx_rows = cb->ctx_rows;
	a.meta = dev ? cb->d_meta : nullptr;
	a.launched_bits = uint32_t(maxb);
	a.prev0_dev = d_prev0;
	a.bit0 = uint32_t(bit_base & 7);
	a.bit_base_dev = d_bit_base;
	a.stage_words = stage_words;
	a.out_words = reinterpret_cast<uint32_t*>(d_out);
	a.out_capacity_words = out_capacity / 4;
	a.agg = reinterpret_cast<unsigned long long*>(ws->enc_desc);
	a.inc = a.agg + tiles;
	a.grp_acc = a.grp_inc = nullptr;
	a.tail = reinterpret_cast<uint32_t*>(a.inc + tiles);
	a.ticket = ws->counters;
	a.result = d_result;
	a.n_tiles = uint32_t(tiles);
	a.full_tiles = uint32_t(n / tile_bytes);
	a.fail = nullptr; a.gate = nullptr; a.alt_result = nullptr;
	const bool aligned = (reinterpret_cast<uint64_t>(d_in) & 15) == 0;
	size_t smem = ((table_bytes + 15) & ~size_t(15)) + (size_t(stage_words) + 8) * sizeof(uint32_t);
	// Context rows + aligned input: whole tiles travel into shared memory by bulk copy (TMA), one tile ahead, when the
	// input buffer still fits next to table and staging area with two CTAs per SM.
	const size_t inbuf_bytes = size_t(kEncThreads) * 32 + 32;
	bool tma = fmt == FMT_CTX && aligned && tunable(kTunEncTma) != 0 && (dev || smem + inbuf_bytes <= size_t(kEncCtxSmemLimit));
	if(tma) smem += inbuf_bytes;
	if(dev) smem = size_t(kEncCtxSmemLimit);
	a.smem_bytes = uint32_t(smem);

	if(warp_tiles) {   // warp-private tiles (K2w)
		const uint64_t wtiles = (n + kWarpTileBytes - 1) / kWarpTileBytes;
		if(wtiles > ws->enc_tiles_cap || wtiles > 0x7fffffffull) return MH_ERR_WORKSPACE;
		const uint64_t wgroups = (wtiles + kEncGroupTiles - 1) / kEncGroupTiles;
		MH_CUDA(cudaMemsetAsync(ws->enc_desc, 0, (wtiles * 2 + wgroups * (kGroupAccStride + 1) + 16) * sizeof(uint64_t), st));   // tile aggregates, tile tails, group prefixes, group accumulators
		a.agg = reinterpret_cast<unsigned long long*>(ws->enc_desc);
		a.inc = a.agg + wtiles;   // the tail words
		a.grp_inc = a.inc + wtiles;
		a.grp_acc = a.grp_inc + ((wgroups + 15) & ~uint64_t(15));   // one accumulator per 128-byte line: 64 additions each, and the scan warp polls them
		a.tail = nullptr;
		a.n_tiles = uint32_t(wtiles);
		a.full_tiles = uint32_t(n / kWarpTileBytes);
		a.stage_words = warp_stage_words(uint32_t(maxb));
		const bool wtma = aligned && tunable(kTunEncTma) != 0;
		const size_t room = size_t(max_smem_optin()) - 2048;   // one CTA per SM (the kernel's static shared memory is ~1.2 KiB): the device-built path takes all of it
		size_t wsmem = dev ? room : size_t(warp_kernel_smem(cb->ctx_rows, a.stage_words, wtma));
		if(wsmem <= room) {
			a.smem_bytes = uint32_t(wsmem);
			auto kern = cb->order ? (wtma ? encode_warp_kernel<1, true> : encode_warp_kernel<1, false>) : (wtma ? encode_warp_kernel<0, true> : encode_warp_kernel<0, false>);
			MH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(wsmem)));
			uint64_t grid = uint64_t(sm_count());
			const uint64_t ctas_needed = (wtiles + kEncWarpThreads / 32 - 1) / (kEncWarpThreads / 32);
			if(grid > ctas_needed) grid = ctas_needed;
			{
				ProfScope p("encode_kernel", st);
				kern<<<unsigned(grid), kEncWarpThreads, wsmem, st>>>(a);
			}
			count_launch(1);
			MH_CUDA(cudaGetLastError());
			return MH_OK;
		}
		// a host-built table too large for one CTA's shared memory: the CTA-wide kernel below
		a.agg = reinterpret_cast<unsigned long long*>(ws->enc_desc);
		a.inc = a.agg + tiles;
		a.tail = reinterpret_cast<uint32_t*>(a.inc + tiles);
		a.n_tiles = uint32_t(tiles);
		a.full_tiles = uint32_t(n / tile_bytes);
		a.stage_words = stage_words;
		MH_CUDA(cudaMemsetAsync(ws->enc_desc, 0, tiles * (2 * sizeof(uint64_t) + sizeof(uint32_t)), st));
	}
	const char* name32 = "encode_kernel";
	if(dev && fmt == FMT_CTX && tma && tunable(kTunEncSpt) != 32) {
		// Device-built tables: the optimistic 64-symbols-per-thread launch first (its own descriptors, ticket and result
		// words), then the ordinary launch, which leaves at once unless the first one declined the tables or failed.
		const uint64_t tile64 = uint64_t(kEncThreads) * 64, tiles64 = (n + tile64 - 1) / tile64;
		uint8_t* base64 = reinterpret_cast<uint8_t*>(ws->enc_desc) + ((tiles * (2 * sizeof(uint64_t) + sizeof(uint32_t)) + 15) & ~uint64_t(15));
		MH_CUDA(cudaMemsetAsync(base64, 0, tiles64 * (2 * sizeof(uint64_t) + sizeof(uint32_t)), st));
		MH_CUDA(cudaMemsetAsync(ws->counters + 2, 0, 2 * sizeof(uint32_t), st));            // the second launch's ticket, the fail flag
		MH_CUDA(cudaMemsetAsync(ws->counters + 8, 0, 4 * sizeof(unsigned long long), st));   // the first launch's result words
		EncArgs b = a;
		b.agg = reinterpret_cast<unsigned long long*>(base64);
		b.inc = b.agg + tiles64;
		b.tail = reinterpret_cast<uint32_t*>(b.inc + tiles64);
		b.result = reinterpret_cast<unsigned long long*>(ws->counters + 8);
		b.n_tiles = uint32_t(tiles64);
		b.full_tiles = uint32_t(n / tile64);
		b.fail = ws->counters + 3;
		const int rc = cb->order ? launch_kernel(encode_kernel<64, FMT_CTX, true, 1, true>, b, smem, st, "encode_kernel")
		                         : launch_kernel(encode_kernel<64, FMT_CTX, true, 0, true>, b, smem, st, "encode_kernel");
		if(rc != MH_OK) return rc;
		a.ticket = ws->counters + 2;
		a.gate = ws->counters + 3;
		a.alt_result = b.result;
		name32 = "encode_tail_kernel";
	}
	if(fmt == FMT_CTX && tma) return cb->order ? launch_variant<32, FMT_CTX, 1, true>(aligned, a, smem, st, name32) : launch_variant<32, FMT_CTX, 0, true>(aligned, a, smem, st, name32);
	if(fmt == FMT_CTX) return cb->order ? launch_variant<32, FMT_CTX, 1>(aligned, a, smem, st) : launch_variant<32, FMT_CTX, 0>(aligned, a, smem, st);
	if(fmt == FMT_WIDE) return launch_variant<16, FMT_WIDE>(aligned, a, smem, st);
	if(fmt == FMT_BOX_SMEM) return spt == 32 ? launch_variant<32, FMT_BOX_SMEM>(aligned, a, smem, st) : launch_variant<16, FMT_BOX_SMEM>(aligned, a, smem, st);
	return spt == 32 ? launch_variant<32, FMT_BOX_GLOBAL>(aligned, a, smem, st) : launch_variant<16, FMT_BOX_GLOBAL>(aligned, a, smem, st);
}

}  // namespace mh
