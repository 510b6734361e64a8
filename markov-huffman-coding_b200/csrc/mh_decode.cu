// mh_decode.cu — kernel 3: self-synchronising speculative parallel decoder.
//
// Replaces the loop of i_coding_provider::decompress (reference src/coding.cpp:118-157) and the bit reader it
// drives (bitbuffer::pop_rest / try_pop_bit / pop_bit, src/bitbuffer.cpp:82-90, :116-140): take the next 8
// stream bits (zero padded past the end), look them up in the previous symbol's 256-entry table
// (decoding_lookup, src/markov_huffman.cpp:56-58); a leaf gives (symbol, length <= 8); an internal node at
// depth 8 consumes 8 bits and is walked bit by bit to a leaf (src/coding.cpp:129-149).
//
// The stream has no symbol count and no block index, so decoding is sequential by construction. Decode state is
// (bit position, previous symbol). The payload is cut into subsequences of S bits; Huffman streams
// self-synchronise, so a decoder started at a wrong state converges onto the true parse.
//   D1 sync     one thread per subsequence decodes it from a guessed state, then keeps decoding its successors
//               until the end state it produces equals the one already recorded there (Weißenberger & Schmidt's
//               scheme, extended to compare the context as well as the bit position). Each CTA also re-decodes
//               the last kDecWarmSubs subsequences of the previous chunk as warm-up, so its own first
//               subsequence normally starts from a converged state.
//   D2 seams    one thread per chunk boundary checks that the previous chunk's recorded end state equals what
//               this chunk's warm-up produced; only on a mismatch does it re-decode from the recorded state until
//               it merges. Repeats until no chunk end state changes (normally zero re-decodes).
//   D3 offsets  per-chunk symbol totals and, for every subsequence, the symbols of its chunk before it; exclusive
//               scan of the totals -> output offset of every subsequence, total output size.
//   D4 write    one thread per subsequence decodes from its verified start state; warps take 32 subsequences at a
//               time from an atomic ticket. The symbols go through per-thread rings in shared memory and leave
//               warp-wide as whole 64-byte units (pair-table path), or with per-thread 8-byte stores.
// D1 and D4 are persistent CTAs that stage the decode table in shared memory once: the two-symbol pair table over the
// live contexts (~50-75 KiB for text, see below) or, with more than 63 live contexts, the reference's 8-bit LUT
// (128 KiB in Markov mode). The payload reaches every thread through a private cp.async ring in shared memory.
#include "mh_internal.hpp"

namespace mh {

namespace {


// LUT entry (u16), see CodingTable::flatten_dectable:
//   leaf : symbol << 8 | length (1..8)           deep : node << 7 | 0x10 (internal node at depth 8)
//   null : ' ' << 8 | 0x20 | 1 (no such table entry: harmless while speculating, an error once verified)
constexpr uint32_t kDeep = 0x10u, kNull = 0x20u, kExt = 0x40u;

// MSB-first bit window over the payload: (hi:lo) holds the bits [pos, loaded) left-aligned, `nextw` is the word
// after them. pos / loaded are bit offsets relative to an origin chosen by the caller.
//
// The payload reaches the window through a small ring in shared memory that every thread owns privately:
// kRingPieces aligned 16-byte chunks, filled by cp.async (global -> shared, no registers, L1 bypassed). With plain
// loads, one word ahead, almost every iteration of a warp had one lane of 32 whose next word missed, and the warp
// waited for it. The ring is topped up on a fixed cadence — every kRefillEvery-th trip of the decode loop, at the same
// trip for all lanes of a warp, with predicated copies — because a refill that each lane triggers when IT
// crosses a chunk boundary is rare per lane but happens in most iterations of the warp, and then the whole warp
// steps through the refill code for a handful of lanes (measured: 78 % of D1's iterations, 6 lanes active).
// A trip consumes at most 32 bits (the slow paths add rounds of their own), so a round sees at most one chunk (four words) used up;
// a chunk is requested two rounds (>= 256 stream bits) before its first word can be popped, and each round waits
// for the copies of the round before it.
// The pieces of the 32 lanes of a warp are interleaved (piece j of lane l at (j * 32 + l) * 16) to spread the banks.
constexpr uint32_t kRingPieces = 4;
constexpr uint32_t kRingBytesPerThread = kRingPieces * 16;
constexpr uint32_t kRefillEvery = 4;

__device__ __forceinline__ void cp_async_16_if(uint32_t dst_shared, const void* src, uint32_t src_bytes, bool go) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}" ::"r"(dst_shared), "l"(src), "r"(src_bytes),
	             "r"(uint32_t(go))
	             : "memory");
}
__device__ __forceinline__ void cp_async_16_full_if(uint32_t dst_shared, uint64_t src, bool go) {   // the whole 16 bytes, no size operand
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst_shared), "l"(src), "r"(uint32_t(go)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Cursor {
	const uint32_t* words;   // payload (4-byte aligned)
	uint64_t n_bytes;        // payload bytes; reads past them return zero (pop_rest pads, src/bitbuffer.cpp:129-140)
	uint64_t base16;         // the 16-byte boundary at or before `words`
	uint32_t ring;           // shared-space address of this thread's piece 0 (set once by the kernel)
	uint32_t cw;             // next chunk to request, counted from base16
	uint32_t rw;             // next word to pop, counted from the same boundary
	uint32_t hi, lo, nextw;  // nextw is kept in memory (little-endian) order and byte-swapped only when consumed
	uint32_t pos, loaded;
	bool interior;           // every chunk this thread can request lies inside the payload: no bounds arithmetic

	__device__ __forceinline__ void attach(uint32_t ring_base_shared) {
		ring = ring_base_shared + ((threadIdx.x >> 5) * (kRingPieces * 32u) + (threadIdx.x & 31u)) * 16u;
		base16 = reinterpret_cast<uint64_t>(words) & ~uint64_t(15);
		interior = false;
	}
	// `last_bit`: no bit at or beyond it (relative to words[0]) is popped by this thread before the next set_reach()
	__device__ __forceinline__ void set_reach(uint64_t last_bit) {
		// chunks are requested up to kRingPieces ahead of the one being read
		interior = ((reinterpret_cast<uint64_t>(words) & 15) + (last_bit >> 3) + 16u * (kRingPieces + 2)) <= (reinterpret_cast<uint64_t>(words) & 15) + n_bytes;
	}
	// chunk c -> slot c % kRingPieces (if `go`); bytes outside the payload are zero-filled by the copy itself
	__device__ __forceinline__ void request_if(uint32_t c, bool go) {
		const uint32_t dst = ring + ((c << 9) & ((kRingPieces - 1) << 9));
		if(interior) {
			cp_async_16_full_if(dst, base16 + (uint64_t(c) << 4), go);
			return;
		}
		const uint64_t addr = reinterpret_cast<uint64_t>(words);
		const int64_t left = int64_t((addr & 15) + n_bytes) - (int64_t(c) << 4);   // payload bytes from the chunk's start on
		const uint32_t n = left >= 16 ? 16u : (left > 0 ? uint32_t(left) : 0u);
		cp_async_16_if(dst, reinterpret_cast<const void*>(n ? base16 + (uint64_t(c) << 4) : base16), n, go);
	}
	// keep kRingPieces chunks requested from the one being read on: at most one is missing per round (<= 4 pops).
	// The decode loops call this on a WARP-UNIFORM cadence (every kRefillEvery-th trip of the loop, all lanes in the same
	// trip, plus once when a loop is entered): a cadence that each lane keeps for itself drifts apart between the lanes
	// (slow paths, checkpoint records), and then almost every trip of the warp steps through this code for a few lanes.
	__device__ __forceinline__ void refill_round() {
		const bool go = int32_t((rw >> 2) + kRingPieces - cw) > 0;
		request_if(cw, go);
		cw += go ? 1u : 0u;
		cp_async_commit();
		cp_async_wait<2>();   // a chunk is first popped three rounds after its request: two rounds may stay in flight
	}
	__device__ __forceinline__ uint32_t pop() {
		uint32_t w;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(ring + ((rw << 7) & ((kRingPieces - 1) << 9)) + ((rw << 2) & 12u)) : "memory");
		++rw;
		return w;
	}
	__device__ __forceinline__ void seek(uint64_t bit, uint32_t rel) {
		const uint64_t abs_bit = bit + ((reinterpret_cast<uint64_t>(words) & 15) << 3);
		const uint64_t w = abs_bit >> 5;
		const uint32_t off = uint32_t(abs_bit & 31);
		cp_async_wait<0>();   // nothing of a previous subsequence may still land in the ring
		rw = uint32_t(w);
		cw = rw >> 2;
#pragma unroll
		for(uint32_t j = 0; j < kRingPieces; ++j) request_if(cw++, true);
		cp_async_commit();
		cp_async_wait<0>();
		const uint32_t w0 = __byte_perm(pop(), 0, 0x0123), w1 = __byte_perm(pop(), 0, 0x0123);
		nextw = pop();
		request_if(cw, (rw >> 2) + kRingPieces != cw);   // the three pops may have finished the first chunk
		cw = (rw >> 2) + kRingPieces;
		cp_async_commit();
		hi = __funnelshift_l(w1, w0, off);
		lo = w1 << off;
		pos = rel;
		loaded = rel + 64 - off;
	}
	__device__ __forceinline__ void take(uint32_t nbits) {   // nbits < 32
		hi = __funnelshift_l(lo, hi, nbits);
		lo <<= nbits;
		pos += nbits;
	}
	__device__ __forceinline__ void take_group(uint32_t nbits) {   // nbits <= 32 (a whole group of four LUT hits)
		hi = __funnelshift_lc(lo, hi, nbits);
		asm("shl.b32 %0, %0, %1;" : "+r"(lo) : "r"(nbits));   // PTX shl clamps: 32 -> 0
		pos += nbits;
	}
	// Called at least once per 32 bits consumed: keeps more than 32 valid bits in the window.
	__device__ __forceinline__ void top_up() {
		const uint32_t avail = loaded - pos;
		if(avail <= 32) {   // all valid bits sit in hi
			const uint32_t w = __byte_perm(nextw, 0, 0x0123);
			hi |= __funnelshift_rc(w, 0u, avail);
			lo = __funnelshift_rc(0u, w, avail);
			loaded += 32;
			nextw = pop();
		}
	}
};

// One symbol: look the top 8 window bits up in the context's row; leaves take the fast path. Returns the symbol and
// advances the cursor. `row_off` is the byte offset of the context's 256 x u16 row (always 0 for ORDER 0).
template <int ORDER, bool LUT_SHARED>
__device__ __forceinline__ uint32_t decode_one(Cursor& cur, uint32_t lut_s, const uint16_t* __restrict__ lut_g,
                                               const uint32_t* __restrict__ walk, uint32_t& row_off, bool& clean) {
	const uint32_t off = row_off + ((cur.hi >> 23) & 0x1feu);
	uint32_t e;
	if(LUT_SHARED) asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(lut_s + off));
	else e = __ldg(reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(lut_g) + off));
	uint32_t sym;
	if(!(e & (kDeep | kNull))) {
		cur.take(e & 15u);
		sym = e >> 8;
	} else if(e & kNull) {
		clean = false;
		cur.take(1);
		sym = ' ';
	} else {
		// codeword longer than 8 bits (src/coding.cpp:129-149): consume the 8 window bits; the next 8 bits index the
		// node's second-level row (ext), which resolves codewords up to 16 bits in one lookup; anything deeper, or a
		// deep node without a row, is walked bit by bit through the flattened tree.
		cur.take(8);
		cur.top_up();   // up to three leaves may precede this symbol in a group: make the next 8 bits valid
		uint32_t node = e >> 7;
		bool walk_it = true;
		sym = ' ';
		if(e & kExt) {
			const uint16_t* ext = reinterpret_cast<const uint16_t*>(walk + (ORDER ? 256 : 1) * 512);
			const uint32_t e2 = __ldg(ext + (node << 8) + (cur.hi >> 24));
			if(e2 & kDeep) {
				cur.take(8);
				node = e2 >> 7;
			} else {
				cur.take(e2 & 15u);
				sym = e2 >> 8;
				walk_it = false;
			}
		}
		if(walk_it) {
			const uint32_t* nodes = walk + row_off;   // row_off = ctx * 512 is also the context's offset in the walk table
			for(int guard = 0; guard < 256; ++guard) {
				cur.top_up();
				if((guard & 15) == 0) cur.refill_round();   // the walk may consume more than a trip's 32 bits: it tops the ring up itself
				const uint32_t bit = cur.hi >> 31;
				cur.take(1);
				const uint32_t w = __ldg(nodes + node);
				const uint32_t child = bit ? (w & 0xffffu) : (w >> 16);
				if(child & 0x8000u) { sym = child & 255u; break; }
				node = child;
				if(guard == 255) clean = false;
			}
		}
		cur.top_up();   // restore the "more than 32 valid bits" invariant the 4-symbol groups rely on
	}
	if(ORDER) row_off = sym << 9;
	return sym;
}

// Four symbols at once on the fast path. The window is left untouched: symbol j peeks at (hi:lo) << consumed,
// where `consumed` is the low bits of `acc`, the running SUM of the LUT entries (a leaf entry carries its length
// in bits [3:0] and nothing in [7:4], so four of them add up to <= 32 in bits [5:0]; the symbol bytes above only
// make junk that the wrap-around funnel shift ignores). The entries are also ORed together: one test after the
// group tells whether any of the four was a deep or null entry, in which case nothing is committed and the caller
// decodes the group one symbol at a time (decode_one). `row` is the shared-space address of the context's row.
// Returns true when the group was committed; `packed` receives the four symbols, first symbol in the low byte.
template <int ORDER>
__device__ __forceinline__ bool decode_four(Cursor& cur, uint32_t lut_s, uint32_t& row, uint32_t& packed) {
	uint32_t acc = 0, flg = 0, r = row, out = 0;
#pragma unroll
	for(int j = 0; j < 4; ++j) {
		const uint32_t t = __funnelshift_l(cur.lo, cur.hi, acc);
		uint32_t e;
		asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(r + ((t >> 24) << 1)));
		acc += e;
		flg |= e;
		out = __byte_perm(out, e, 0x5321);   // shift in byte 1 of e (the symbol)
		if(ORDER) r = lut_s + ((e & 0xff00u) << 1);
	}
	if(flg & (kDeep | kNull)) return false;
	cur.take_group(acc & 63u);
	row = r;
	packed = out;
	return true;
}

// Decode every symbol whose first bit lies in [cur.pos, limit), counting them. The window is topped up once per
// group of four symbols (four LUT hits consume at most 32 bits), so lanes refill together instead of one by one.
template <int ORDER, bool LUT_SHARED>
__device__ __forceinline__ bool decode_until(Cursor& cur, uint32_t lut_s, const uint16_t* __restrict__ lut_g,
                                             const uint32_t* __restrict__ walk, uint32_t limit, uint32_t& ctx, uint32_t& count) {
	bool clean = true;
	uint32_t sym = ctx;
	uint32_t row_off = ORDER ? ctx << 9 : 0u;
	uint32_t trip = 0;
	cur.refill_round();
	while(cur.pos < limit) {
		if((++trip & 1) == 0) cur.refill_round();   // four symbols of up to 16 bits per trip (longer ones walk, and top up themselves)
#pragma unroll
		for(int j = 0; j < 4; ++j) {
			if(cur.pos < limit) {
				sym = decode_one<ORDER, LUT_SHARED>(cur, lut_s, lut_g, walk, row_off, clean);
				++count;
			}
		}
		cur.top_up();
	}
	ctx = sym;
	return clean;
}

// D1's walker: decode one subsequence of kCp segments in ONE flat loop, so the lanes of a warp only wait for each
// other at the end of the subsequence, not at every checkpoint. A trip decodes four symbols speculatively (as
// decode_four) and commits all of them while the group ends before the current checkpoint `cp_end` and holds no deep
// or null entry — the only test on the main path. Otherwise ONE divergent block finishes the job in the same trip: it
// commits the leading entries that end before the checkpoint, then exactly one more symbol (through the full path if
// its entry is flagged). If that symbol ends at or after the checkpoint the walker now rests on the first codeword
// boundary at or after it — the state a checkpoint records, with the symbols that started in the segment. (Until
// round 2 a group that straddled a checkpoint was cut and the walker crept up to the boundary over several trips:
// rare per lane, but with 32 lanes and 8 checkpoints a third of the warp's trips stepped through those paths.)
// With `compare`, the walk stops at the first checkpoint whose recorded state equals the walker's (the recorded
// trajectory is then its own). `cp_st` / `cp_cn` point at this subsequence's column of the [kCp][kDecThreads]
// checkpoint arrays. Returns true when it stopped on a match.
constexpr int kCp = 8;

__device__ __forceinline__ uint32_t pack_cp(uint32_t rel_bits, uint32_t ctx) { return ((rel_bits > 255u ? 255u : rel_bits) << 8) | ctx; }

// The walker rests on a codeword boundary at `at` (relative to the CTA's origin): record every checkpoint at or before
// it. 0: go on, 1: a recorded state equalled the walker's, 2: that was the subsequence's last checkpoint.
struct CpWalk {
	uint32_t sub_begin, unit, span, cp_end;
	uint64_t cp_tab;
	int j;
	__device__ __forceinline__ void start(uint32_t sub_begin_, uint32_t unit_, uint64_t cp_tab_, uint32_t span_) {
		sub_begin = sub_begin_; unit = unit_; cp_tab = cp_tab_; span = span_;
		j = 0;
		set_end();
	}
	__device__ __forceinline__ void set_end() {   // checkpoint j, in units of 1/32 subsequence
		cp_end = sub_begin + (uint32_t(cp_tab >> (8 * j)) & 255u) * unit;
		if(cp_end > span) cp_end = span;
	}
	__device__ __forceinline__ int cross(uint32_t at, uint32_t ctx_or_row, uint32_t& cnt, uint16_t* cp_st, uint16_t* cp_cn, bool compare) {
		while(at >= cp_end) {
			const uint16_t st = uint16_t(pack_cp(at - cp_end, ctx_or_row));
			cp_cn[j * kDecThreads] = uint16_t(cnt);
			if(compare && cp_st[j * kDecThreads] == st) return 1;
			cp_st[j * kDecThreads] = st;
			cnt = 0;
			if(++j == kCp) return 2;
			set_end();
		}
		return 0;
	}
};

template <int ORDER>
__device__ __forceinline__ bool walk_subsequence(Cursor& cur, uint32_t lut_s, const uint16_t* __restrict__ lut_g,
                                                 const uint32_t* __restrict__ walk, uint32_t sub_begin, uint32_t unit, uint64_t cp_tab, uint32_t span,
                                                 uint32_t& row, uint16_t* cp_st, uint16_t* cp_cn, bool compare) {
	uint32_t cnt = 0, trip = 0;
	bool clean = true;
	CpWalk cw;
	cw.start(sub_begin, unit, cp_tab, span);
	cur.refill_round();
	if(cur.pos >= cw.cp_end) {   // a walker that came out of the previous subsequence beyond this one's first checkpoint (tiny subsequences, long codewords)
		const int s = cw.cross(cur.pos, ORDER ? (row - lut_s) >> 9 : 0u, cnt, cp_st, cp_cn, compare);
		if(s) return s == 1;
	}
	for(;;) {
		if((++trip & (kRefillEvery - 1)) == 0) cur.refill_round();   // every lane of the warp in the same trip
		// four speculative LUT hits
		uint32_t e[4], a[5], r[5], flg = 0;
		a[0] = 0; r[0] = row;
#pragma unroll
		for(int i = 0; i < 4; ++i) {
			const uint32_t t = __funnelshift_l(cur.lo, cur.hi, a[i]);
			asm("ld.shared.u16 %0, [%1];" : "=r"(e[i]) : "r"(r[i] + ((t >> 24) << 1)));
			a[i + 1] = a[i] + e[i];
			flg |= e[i];
			r[i + 1] = ORDER ? lut_s + ((e[i] & 0xff00u) << 1) : row;
		}
		const uint32_t room = cw.cp_end - cur.pos;   // > 0
		if(!(flg & (kDeep | kNull)) && (a[4] & 63u) < room) {   // the whole group ends before the checkpoint
			cur.take_group(a[4] & 63u);
			row = r[4];
			cnt += 4;
			cur.top_up();
			continue;
		}
		// the leading entries that end before the checkpoint, then one symbol more
		const bool c0 = !(e[0] & (kDeep | kNull)) && (a[1] & 63u) < room;
		const bool c1 = c0 && !(e[1] & (kDeep | kNull)) && (a[2] & 63u) < room;
		// (a trip may consume 32 bits at most — the ring's refill cadence counts on it: three entries are committed only
		// before an ordinary fourth one; a flagged one, up to 16 bits without the walk, waits for the next trip)
		const bool c2 = c1 && !((e[2] | e[3]) & (kDeep | kNull)) && (a[3] & 63u) < room;
		const uint32_t n = uint32_t(c0) + uint32_t(c1) + uint32_t(c2);
		const uint32_t x = c2 ? e[3] : (c1 ? e[2] : (c0 ? e[1] : e[0]));
		cur.take_group((c2 ? a[3] : (c1 ? a[2] : (c0 ? a[1] : 0u))) & 63u);
		row = c2 ? r[3] : (c1 ? r[2] : (c0 ? r[1] : r[0]));
		cnt += n + 1;
		if(x & (kDeep | kNull)) {   // rare: through the full path
			uint32_t row_off = row - lut_s;
			decode_one<ORDER, true>(cur, lut_s, lut_g, walk, row_off, clean);
			row = lut_s + row_off;
		} else {
			cur.take(x & 15u);
			if(ORDER) row = lut_s + ((x & 0xff00u) << 1);
		}
		cur.top_up();   // before a return: the caller walks on into the next subsequence with this window
		const int s = cw.cross(cur.pos, ORDER ? (row - lut_s) >> 9 : 0u, cnt, cp_st, cp_cn, compare);
		if(s) return s == 1;
	}
}

// ---------------------------------------------------------------------------------------------------------
// Pair table (CodingTable::flatten_pairlut): u32 entries over the LIVE contexts only, each resolving up to two
// symbols from the next 8 stream bits. Entry: [5:0] bits consumed (bits 4, 5 = deep / null flags), [9:6] symbols
// produced, [15:10] next row, [23:16] first symbol, [31:24] second symbol or 0. Four entries are summed like the u16
// entries above: the low 6 bits of the sum are the bits consumed (<= 32), bits [9:6] the symbols produced (<= 8).
// A codeword of 9..16 bits is two ordinary entries: a prefix entry (8 bits, no symbol, next row = the prefix row of
// its depth-8 node) and an entry of that prefix row — no branch. Text has a few dozen live contexts, so the table
// takes ~50 KiB of shared memory instead of 128 KiB, and one dependent lookup yields ~1.85 symbols. Flagged entries
// (longer codewords, missing table entries) go through the u16 LUT / walk table in global memory (decode_one).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kPairFlags = kDeep | kNull;
constexpr uint32_t kPairCount = 0xc0u;   // symbols produced by one entry: 0x40 one, 0x80 two, 0 = prefix entry

// Rows are named by their byte offset inside the table (row r = r * 1024): the entry's next-row field [15:10] already
// is that offset, so a lookup address is (entry & 0xfc00) | (window byte * 4) — one LOP3 — plus the table's base,
// which rides in the load's uniform-register operand.
struct PairTab {
	uint32_t tab;    // shared-space address of the table (row r at tab + r * 1024)
	uint32_t rank;   // shared-space address of rank[256]: row of a byte as context
	uint32_t live;   // shared-space address of live[64]: context byte of a context row
	uint32_t len1;   // shared-space address of len1[ctx_rows * 256]: length of an entry's first codeword
	uint32_t null_row;   // offset of the null row
};

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
	uint32_t v;
	asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

__host__ __device__ __forceinline__ uint32_t pair_table_bytes(uint32_t rows, uint32_t ctx_rows) { return rows * 1024u + 256u + 64u + ctx_rows * 256u; }

// Stage the pair table and its byte maps in shared memory (all threads; the caller synchronises).
__device__ __forceinline__ PairTab stage_pair_table(uint32_t* smem, const uint32_t* __restrict__ pair_g, uint32_t rows, uint32_t ctx_rows) {
	const uint32_t n16 = pair_table_bytes(rows, ctx_rows) / 16u;
	const uint4* src = reinterpret_cast<const uint4*>(pair_g);
	uint4* dst = reinterpret_cast<uint4*>(smem);
	for(uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
	PairTab T;
	asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(T.tab) : "l"(smem));
	T.rank = T.tab + rows * 1024u;
	T.live = T.rank + 256u;
	T.len1 = T.live + 64u;
	T.null_row = ctx_rows * 1024u;
	return T;
}

template <int ORDER>
__device__ __forceinline__ uint32_t pair_row_of(const PairTab& T, uint32_t ctx) { return ORDER ? lds_u8(T.rank + ctx) << 10 : 0u; }

// Four speculative lookups; the window is left untouched. a[i] = sum of the first i entries, r[i] = row after them.
template <int ORDER>
__device__ __forceinline__ void pair_lookups(const Cursor& cur, const PairTab& T, uint32_t row, uint32_t (&e)[4], uint32_t (&a)[5], uint32_t (&r)[5]) {
	a[0] = 0;
	r[0] = row;
#pragma unroll
	for(int i = 0; i < 4; ++i) {
		const uint32_t t = __funnelshift_l(cur.lo, cur.hi, a[i]);
		e[i] = lds_u32(T.tab + (r[i] | ((t >> 22) & 0x3fcu)));
		a[i + 1] = a[i] + e[i];
		r[i + 1] = e[i] & 0xfc00u;   // order 0: context row 0, or a prefix row
	}
}

// One symbol through the reference's path (u16 LUT + walk table in global memory), keeping `row` in step.
// `row` is a context row or the null row — or a prefix row, when the window selects a deep flag there: a codeword of 17
// bits or more, of which the walker has taken the first 8; the flag names the tree node its next 8 bits lead to, and the
// context whose tree that is, and the walk goes on from there bit by bit (src/coding.cpp:137-149).
template <int ORDER>
__device__ __forceinline__ uint32_t pair_slow_one(Cursor& cur, const PairTab& T, const uint16_t* __restrict__ lut_g,
                                                  const uint32_t* __restrict__ walk, uint32_t& row, bool& clean) {
	if(row > T.null_row) {
		const uint32_t x = lds_u32(T.tab + (row | ((cur.hi >> 22) & 0x3fcu)));
		cur.take(8);
		const uint32_t* nodes = walk + (ORDER ? ((x >> 16) & 255u) << 9 : 0u);
		uint32_t node = (x >> 24) | ((x & 0x40u) << 2), sym = ' ';
		for(int guard = 0; guard < 256; ++guard) {
			cur.top_up();
			if((guard & 15) == 0) cur.refill_round();   // the walk tops the ring up itself
			const uint32_t bit = cur.hi >> 31;
			cur.take(1);
			const uint32_t w = __ldg(nodes + node);
			const uint32_t child = bit ? (w & 0xffffu) : (w >> 16);
			if(child & 0x8000u) { sym = child & 255u; break; }
			node = child;
			if(guard == 255) clean = false;
		}
		cur.top_up();
		if(ORDER) row = lds_u8(T.rank + sym) << 10;
		else row = 0u;
		return sym;
	}
	uint32_t row_off = ORDER ? lds_u8(T.live + (row >> 10)) << 9 : 0u;
	const uint32_t sym = decode_one<ORDER, false>(cur, 0u, lut_g, walk, row_off, clean);
	if(ORDER) row = lds_u8(T.rank + sym) << 10;
	return sym;
}

// Exactly one symbol, given the entry e0 the window selects in `row`: a single-symbol entry is taken as it is, a
// pair gives up its first symbol only (its length comes from len1), a prefix or flagged entry goes through the global
// tables (a flagged entry of a prefix row: from the node it names).
template <int ORDER>
__device__ __forceinline__ uint32_t pair_step_one(Cursor& cur, const PairTab& T, const uint16_t* __restrict__ lut_g,
                                                  const uint32_t* __restrict__ walk, uint32_t e0, uint32_t& row, bool& clean) {
	if((e0 & kPairFlags) || !(e0 & kPairCount)) return pair_slow_one<ORDER>(cur, T, lut_g, walk, row, clean);
	const uint32_t sym = (e0 >> 16) & 255u;
	if(e0 & 0x80u) {
		cur.take(lds_u8(T.len1 + (row >> 2) + (cur.hi >> 24)));
		row = pair_row_of<ORDER>(T, sym);
	} else {
		cur.take(e0 & 15u);
		row = e0 & 0xfc00u;
	}
	return sym;
}

// walk_subsequence over the pair table. The main path commits a whole group that ends before the checkpoint (it may end
// on a prefix entry: the walker then rests inside a codeword, in its prefix row, and the next trip completes it). The
// divergent block commits the leading entries that end before the checkpoint and then exactly one codeword: a flagged
// entry through the global tables, a prefix entry together with its completion, a single entry as it is, and of a pair
// both symbols only if the first one still ends before the checkpoint (its length comes from len1) — so the walker rests
// on the first codeword boundary at or after the checkpoint whenever it has passed it.
// Checkpoints record the ROW (not the context byte): equality is all the comparison needs.
template <int ORDER>
__device__ __forceinline__ bool walk_subsequence_pair(Cursor& cur, const PairTab& T, const uint16_t* __restrict__ lut_g,
                                                      const uint32_t* __restrict__ walk, uint32_t sub_begin, uint32_t unit, uint64_t cp_tab, uint32_t span,
                                                      uint32_t& row, uint16_t* cp_st, uint16_t* cp_cn, bool compare) {
	uint32_t cnt = 0;
	bool clean = true;
	CpWalk cw;
	cw.start(sub_begin, unit, cp_tab, span);
	cur.refill_round();
	if(cur.pos >= cw.cp_end) {   // a walker that came out of the previous subsequence beyond this one's first checkpoint (tiny subsequences, long codewords)
		const int s = cw.cross(cur.pos, row >> 10, cnt, cp_st, cp_cn, compare);
		if(s) return s == 1;
	}
	// one trip of the walk: 0 to go on, else cw.cross()'s verdict (1: the recorded state was met, 2: the subsequence ended)
	auto step = [&]() -> int {
		uint32_t e[4], a[5], r[5];
		pair_lookups<ORDER>(cur, T, row, e, a, r);
		const uint32_t room = cw.cp_end - cur.pos;   // > 0
		if(!((e[0] | e[1] | e[2] | e[3]) & kPairFlags) && (a[4] & 63u) < room) {
			cur.take_group(a[4] & 63u);
			row = r[4];
			cnt += (a[4] >> 6) & 15u;
			cur.top_up();
			return 0;
		}
		const bool c0 = !(e[0] & kPairFlags) && (a[1] & 63u) < room;
		const bool c1 = c0 && !(e[1] & kPairFlags) && (a[2] & 63u) < room;
		// (a trip may consume 32 bits at most — the ring's refill cadence counts on it: three entries are committed only
		// before an ordinary fourth one; a flagged or prefix entry, up to 16 bits, waits for the next trip)
		const bool c2 = c1 && !((e[2] | e[3]) & kPairFlags) && (e[3] & kPairCount) && (a[3] & 63u) < room;
		const uint32_t an = c2 ? a[3] : (c1 ? a[2] : (c0 ? a[1] : 0u));
		const uint32_t x = c2 ? e[3] : (c1 ? e[2] : (c0 ? e[1] : e[0]));
		const uint32_t x1 = c1 ? e[3] : (c0 ? e[2] : e[1]);   // the entry after x (unless x is the group's last)
		cur.take_group(an & 63u);
		row = c2 ? r[3] : (c1 ? r[2] : (c0 ? r[1] : r[0]));
		cnt += (an >> 6) & 15u;
		const uint32_t left = room - (an & 63u);   // bits from here to the checkpoint (> 0)
		if(x & kPairFlags) {
			pair_slow_one<ORDER>(cur, T, lut_g, walk, row, clean);
			++cnt;
		} else if(!(x & kPairCount)) {   // the first 8 bits of a longer codeword: it ends with the next entry (one symbol of the prefix row)
			if(!c2) {
				if(x1 & kPairFlags) {   // 17 bits or more: the prefix row names the node, the walk goes on from there
					cur.take(8);
					row = x & 0xfc00u;
					pair_slow_one<ORDER>(cur, T, lut_g, walk, row, clean);
				} else {
					cur.take_group(8u + (x1 & 15u));
					row = x1 & 0xfc00u;
				}
				++cnt;
			}   // else: the group's last entry — the next trip starts with it
		} else if(x & 0x80u) {   // two symbols
			const uint32_t l1 = lds_u8(T.len1 + (row >> 2) + (cur.hi >> 24));
			if(l1 >= left) {   // the first one already reaches the checkpoint
				cur.take(l1);
				row = pair_row_of<ORDER>(T, (x >> 16) & 255u);
				++cnt;
			} else {
				cur.take(x & 15u);
				row = x & 0xfc00u;
				cnt += 2;
			}
		} else {
			cur.take(x & 15u);
			row = x & 0xfc00u;
			++cnt;
		}
		cur.top_up();   // before a return: the caller walks on into the next subsequence with this window
		return cw.cross(cur.pos, row >> 10, cnt, cp_st, cp_cn, compare);
	};
	// kRefillEvery trips in a row, the ring's refill at a fixed place among them (every lane of the warp in the same trip):
	// no trip counter and no cadence test in the loop
	for(;;) {
#pragma unroll
		for(uint32_t u = 0; u < kRefillEvery; ++u) {
			if(u == kRefillEvery - 1) cur.refill_round();
			const int s = step();
			if(s) return s == 1;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// D4 output staging. A thread's symbols are consecutive in memory, but the 32 threads of a warp write 32 different
// places: as plain stores that is 32 memory transactions per warp instruction, and those transactions — not the
// decoding — were what D4 spent a third of its time on. So the symbols go through shared memory: every thread owns
// a 128-byte ring that mirrors global memory (ring byte = address & 127), i.e. two 64-byte units; completed words
// are stored into it, and every few iterations the warp writes out the units that are complete, four lanes per
// unit with one 16-byte store each: whole sectors, eight units per warp instruction. The first and last unit of a
// thread's range are written with byte masks.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kOutRingBytes = 128;
constexpr uint32_t kFlushEvery = 8;   // iterations between flush rounds. At a round a thread's current unit holds < 64 bytes and
                                      // everything before it has left; 8 iterations add <= 64 bytes, so the thread cannot reach the
                                      // ring half of that unit again before the next round has written it out

__device__ __forceinline__ void sts_u32_if(uint32_t addr, uint32_t v, bool go) {
	if(go) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");   // the compiler predicates the store on the comparison itself
}

// 16 bytes of a unit to global memory; only bytes [lo, hi) of the piece (0 <= lo, hi <= 16 after clamping) are written
__device__ __forceinline__ void store_piece(uint8_t* dst, const uint4& v, int lo, int hi) {
	if(lo <= 0 && hi >= 16) {
		__stcg(reinterpret_cast<uint4*>(dst), v);
		return;
	}
	const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
	for(int k = 0; k < 16; ++k)
		if(k >= lo && k < hi) dst[k] = uint8_t(w[k >> 2] >> (8 * (k & 3)));
}

struct OutStage {
	uint32_t ring;        // shared-space address of this thread's ring
	uint32_t ring_warp;   // shared-space address of lane 0's ring
	uint64_t unit;        // global address of the first unit not yet written out; bits [5:0]: first valid byte in it
	uint32_t g;           // low 32 bits of the global address of the next symbol
	uint32_t pend;        // the bytes of the word at g & ~3 produced so far (g & 3 of them)

	__device__ __forceinline__ void begin(uint8_t* out) {
		const uint64_t a = reinterpret_cast<uint64_t>(out);
		unit = a;            // (a & ~63) | (a & 63)
		g = uint32_t(a);
		pend = 0;
	}
	// append cnt (<= 8) symbol bytes, first symbol in the low byte of lo
	__device__ __forceinline__ void append(uint32_t lo, uint32_t hi, uint32_t cnt) {
		const uint32_t s = (g & 3u) * 8u;
		const uint32_t t0 = lo << s, t1 = __funnelshift_l(lo, hi, s), t2 = __funnelshift_l(hi, 0u, s);
		const uint32_t x0 = pend | t0;
		const uint32_t tb = (g & 3u) + cnt;   // <= 11 bytes: up to two complete words
		sts_u32_if(ring + (g & 124u), x0, tb >= 4u);
		sts_u32_if(ring + ((g + 4u) & 124u), t1, tb >= 8u);
		pend = tb >= 8u ? t2 : (tb >= 4u ? t1 : x0);
		g += cnt;
	}
	// Warp-wide: write out every lane's complete unit (tail == false), or what is left of its last unit (tail == true).
	__device__ __forceinline__ void flush(bool tail) {
		const uint32_t lane = threadIdx.x & 31u;
		__syncwarp();
		const uint32_t u32 = uint32_t(unit);
		// complete: the next symbol lies beyond the unit; tail: some byte of the unit has been produced
		const bool ready = tail ? (g != u32) : ((g >> 6) != (u32 >> 6));
		const uint32_t mask = __ballot_sync(0xffffffffu, ready);
		if(mask) {
#pragma unroll
			for(uint32_t j = 0; j < 4; ++j) {
				if(!(mask & (0x11111111u << j))) continue;
				const uint32_t src = (lane & ~3u) + j;
				const uint32_t a_lo = __shfl_sync(0xffffffffu, uint32_t(unit), src);
				const uint32_t a_hi = __shfl_sync(0xffffffffu, uint32_t(unit >> 32), src);
				const uint32_t s_g = __shfl_sync(0xffffffffu, g, src);
				if((mask >> src) & 1u) {
					const uint32_t piece = (lane & 3u) * 16u;
					uint4 v;
					asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring_warp + src * kOutRingBytes + (a_lo & 64u) + piece) : "memory");
					uint8_t* dst = reinterpret_cast<uint8_t*>(((uint64_t(a_hi) << 32) | (a_lo & ~63u)) + piece);
					const int first = int(a_lo & 63u) - int(piece);
					const int last = tail ? int(s_g - (a_lo & ~63u)) - int(piece) : 16;   // tail: the unit ends at the next symbol's address
					store_piece(dst, v, first, last);
				}
			}
			if(ready) unit = (unit & ~uint64_t(63)) + 64;
		}
		__syncwarp();
	}
	// after the last symbol: the incomplete word, the complete unit (if any), the rest
	__device__ __forceinline__ void finish() {
		sts_u32_if(ring + (g & 124u), pend, (g & 3u) != 0);
		flush(false);
		flush(true);
	}
};

// decode_emit over the pair table, warp-collective: every lane of the warp calls it (count == 0: nothing to decode)
// and decodes exactly `count` symbols to `out`. A committed group is 2..8 symbols: the symbol bytes of its four
// entries are compacted with two byte permutes and a shift and appended to the thread's output ring. The last group
// is simply cut off at `count` symbols, unless the exact end position is wanted (`exact_end`: the stream's last
// subsequence), which steps symbol by symbol.
template <int ORDER>
__device__ __forceinline__ bool decode_emit_pair(Cursor& cur, const PairTab& T, const uint16_t* __restrict__ lut_g,
                                                 const uint32_t* __restrict__ walk, uint32_t count, uint32_t ctx, uint8_t* out, bool exact_end,
                                                 OutStage& os) {
	bool clean = true;
	uint32_t row = pair_row_of<ORDER>(T, ctx);
	uint32_t rem = count;
	os.begin(out);
	if(rem) cur.refill_round();
	// one trip of one lane; `tail`: the lane may be within a group of its last symbol (cut the group at `rem`, or step
	// symbol by symbol when the exact end position is wanted)
	auto trip = [&](uint32_t it, bool tail) {
		if((it & (kRefillEvery - 1)) == kRefillEvery - 1) cur.refill_round();   // every lane of the warp in the same trip
		uint32_t e[4], a[5], r[5];
		pair_lookups<ORDER>(cur, T, row, e, a, r);
		const uint32_t f4 = e[0] | e[1] | e[2] | e[3];
		uint32_t gc = (a[4] >> 6) & 15u;
		uint32_t g_lo, g_hi;
		if(!(f4 & kPairFlags) && !(tail && exact_end && gc > rem)) {
			// selector by the number of symbols in the first entry of each half: 0 -> 0x3376, 1 -> 0x3762, 2 -> 0x7632
			const uint32_t sel_a = __funnelshift_rc(0x37623376u, 0x7632u, (e[0] & kPairCount) >> 2);
			const uint32_t sel_b = __funnelshift_rc(0x37623376u, 0x7632u, (e[2] & kPairCount) >> 2);
			const uint32_t wa = __byte_perm(e[0], e[1], sel_a);
			const uint32_t wb = __byte_perm(e[2], e[3], sel_b);
			const uint32_t ca8 = (a[2] >> 3) & 0x78u;   // 8 x symbols of the first two entries (<= 32)
			uint32_t up;
			asm("shl.b32 %0, %1, %2;" : "=r"(up) : "r"(wb), "r"(ca8));   // clamps: 32 -> 0
			g_lo = wa | up;
			g_hi = __funnelshift_lc(wb, 0u, ca8);
			if(tail) gc = gc < rem ? gc : rem;
			cur.take_group(a[4] & 63u);
			row = r[4];
		} else {
			g_lo = pair_step_one<ORDER>(cur, T, lut_g, walk, e[0], row, clean);
			g_hi = 0;
			gc = 1;
		}
		rem -= gc;
		os.append(g_lo, g_hi, gc);
		cur.top_up();
	};
	uint32_t it = 0;
	// (eight trips in a row with the refills and the flush at fixed places and one vote per eight — what paid in D1's walk —
	// measured 6 % slower here, 26 % on the Fibonacci stream: 56 registers instead of 48 and an 800-instruction loop body)
	// while every lane of the warp has at least a whole group (8 symbols) to go: no lane test, no cut, no end check
	for(; __all_sync(0xffffffffu, rem >= 8u); ++it) {
		trip(it, false);
		if((it & (kFlushEvery - 1)) == kFlushEvery - 1) os.flush(false);
	}
	for(; __any_sync(0xffffffffu, rem != 0); ++it) {
		if(rem) trip(it, true);
		if((it & (kFlushEvery - 1)) == kFlushEvery - 1) os.flush(false);
	}
	os.finish();
	return clean;
}

// Decode exactly `count` symbols and write them to `out`. Single bytes until the address is 8-byte aligned, then
// eight symbols per aligned 64-bit store — every lane stores in the same iteration — then the ragged tail.
template <int ORDER>
__device__ __forceinline__ bool decode_emit(Cursor& cur, uint32_t lut_s, const uint16_t* __restrict__ lut_g,
                                            const uint32_t* __restrict__ walk, uint32_t count, uint32_t ctx, uint8_t* out) {
	bool clean = true;
	uint32_t row_off = ORDER ? ctx << 9 : 0u;
	uint32_t head = uint32_t((8 - (reinterpret_cast<uint64_t>(out) & 7)) & 7);
	if(head > count) head = count;
	uint32_t trip = 0;
	cur.refill_round();
	for(uint32_t i = 0; i < head; ++i) {
		*out++ = uint8_t(decode_one<ORDER, true>(cur, lut_s, lut_g, walk, row_off, clean));
		cur.top_up();
		if((++trip & (kRefillEvery - 1)) == 0) cur.refill_round();
	}
	const uint32_t groups = (count - head) >> 3;
	uint32_t row = lut_s + row_off;
	for(uint32_t g = 0; g < groups; ++g) {
		uint32_t w[2];
#pragma unroll
		for(int h = 0; h < 2; ++h) {
			if(!decode_four<ORDER>(cur, lut_s, row, w[h])) {   // a deep or null entry among the four: one symbol at a time
				row_off = row - lut_s;
				w[h] = 0;
#pragma unroll 1
				for(int j = 0; j < 4; ++j) w[h] = __byte_perm(w[h], decode_one<ORDER, true>(cur, lut_s, lut_g, walk, row_off, clean), 0x4321);
				row = lut_s + row_off;
			}
			cur.top_up();
			if((++trip & 1) == 0) cur.refill_round();   // a half is four LUT hits (<= 32 bits) or four symbols of up to 16 bits
		}
		__stcg(reinterpret_cast<uint2*>(out), make_uint2(w[0], w[1]));   // L2 only: the output must not push payload lines out of L1
		out += 8;
	}
	row_off = row - lut_s;
	const uint32_t tail = (count - head) & 7;
	for(uint32_t i = 0; i < tail; ++i) {
		*out++ = uint8_t(decode_one<ORDER, true>(cur, lut_s, lut_g, walk, row_off, clean));
		cur.top_up();
		if((++trip & (kRefillEvery - 1)) == 0) cur.refill_round();
	}
	return clean;
}

// ---------------------------------------------------------------------------------------------------------
// D1: speculative decode + intra-chunk synchronisation
//
// Each subsequence carries kCp checkpoints (at 1, 2, 4, 8, 12, 16, 24, 32 thirty-seconds of its length). Checkpoint j records the decoder state at
// the first codeword boundary at or after it, and how many symbols started in the segment before it. A thread
// that takes over its successor's subsequence stops at the first checkpoint where its own state equals the
// recorded one: from there on the recorded trajectory is its own. Per-segment counts (not running totals) make
// the records of a partially overwritten subsequence consistent whoever wrote which segment.
// ---------------------------------------------------------------------------------------------------------
template <int ORDER, bool PAIR>
__global__ void __launch_bounds__(kDecThreads, 1) dec_sync_kernel(
    const uint32_t* __restrict__ words, uint64_t n_bits, uint64_t buf_bytes, uint32_t start0, const uint16_t* __restrict__ lut_g,
    const uint32_t* __restrict__ walk, const uint32_t* __restrict__ pair_g, uint32_t pair_rows, uint32_t pair_ctx_rows,
    uint32_t* __restrict__ state, uint32_t* __restrict__ count, uint32_t* __restrict__ seam, uint32_t sub_bits, uint64_t n_subs, uint32_t n_chunks, uint32_t warm, uint32_t lut_smem_bytes,
    uint64_t cp_tab) {
	extern __shared__ __align__(16) uint16_t lut_s[];   // table, then the payload rings (kRingBytesPerThread each)
	__shared__ uint16_t cp_state[kCp][kDecThreads];   // [checkpoint][slot]: conflict-free across a warp
	__shared__ uint16_t cp_count[kCp][kDecThreads];
	const uint32_t tid = threadIdx.x;
	const uint32_t chunk_subs = blockDim.x - warm;   // launched with up to kDecThreads threads (fewer for streams with few subsequences)
	const uint32_t unit = sub_bits / 32;
	PairTab T = {};
	if(PAIR) {
		T = stage_pair_table(reinterpret_cast<uint32_t*>(lut_s), pair_g, pair_rows, pair_ctx_rows);
	} else {
		const uint32_t n16 = ORDER ? 65536u : 256u;
		const uint4* src = reinterpret_cast<const uint4*>(lut_g);
		uint4* dst = reinterpret_cast<uint4*>(lut_s);
		for(uint32_t i = tid; i < n16 / 8; i += blockDim.x) dst[i] = src[i];
	}
	__syncthreads();
	uint32_t lut_sa;
	asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(lut_sa) : "l"(lut_s));
	// start of a speculative decode: the context ' ' (or the first live context when ' ' has no tree)
	uint32_t guess_row = lut_sa + (ORDER ? uint32_t(' ') << 9 : 0u);
	if(PAIR) {
		guess_row = pair_row_of<ORDER>(T, ' ');
		if(ORDER && guess_row == T.null_row) guess_row = 0u;   // row 0: the first live context
	}
	Cursor cur;
	cur.words = words;
	cur.n_bytes = buf_bytes;
	cur.attach(lut_sa + lut_smem_bytes);

	for(uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
		// thread t handles subsequence first_sub + t where first_sub may be negative for chunk 0
		const int64_t first_sub = int64_t(chunk) * chunk_subs - warm;
		const int64_t my_sub = first_sub + tid;
		const int64_t end_sub = int64_t(chunk + 1) * chunk_subs < int64_t(n_subs) ? int64_t(chunk + 1) * chunk_subs : int64_t(n_subs);
		// bit origin of this CTA's window; the subsequence grid starts at the stream's first bit (start0 >> 8)
		const uint64_t origin = (start0 >> 8) + (first_sub < 0 ? 0 : uint64_t(first_sub) * sub_bits);
		const int64_t origin_sub = first_sub < 0 ? 0 : first_sub;
		const uint64_t span = n_bits - origin;   // bits from the origin to the end of the stream
		bool active = my_sub >= 0 && my_sub < end_sub;

		const uint32_t span32 = span > 0xffff0000ull ? 0xffff0000u : uint32_t(span);   // a CTA's window is a few Mbit
		uint32_t row = guess_row;
		int64_t k = my_sub;
		if(active) {
			const uint32_t pos = uint32_t(uint64_t(my_sub - origin_sub) * sub_bits);
			if(ORDER && my_sub == 0) row = PAIR ? pair_row_of<ORDER>(T, start0 & 255u) : lut_sa + ((start0 & 255u) << 9);   // the stream's own start (exact, or a shard's guess)
			cur.set_reach(origin + uint64_t(end_sub - origin_sub) * sub_bits + 64);   // a thread walks at most to the end of its chunk
			cur.seek(origin + pos, pos);
			if(PAIR) walk_subsequence_pair<ORDER>(cur, T, lut_g, walk, pos, unit, cp_tab, span32, row, &cp_state[0][tid], &cp_count[0][tid], false);
			else walk_subsequence<ORDER>(cur, lut_sa, lut_g, walk, pos, unit, cp_tab, span32, row, &cp_state[0][tid], &cp_count[0][tid], false);
		}
		__syncthreads();
		// rounds: walk the successor's checkpoints until my state equals the recorded one
		for(;;) {
			++k;
			if(active && k >= end_sub) active = false;
			if(active) {
				const uint32_t slot = uint32_t(k - first_sub);
				const uint32_t begin = uint32_t(uint64_t(k - origin_sub) * sub_bits);
				const bool hit = PAIR ? walk_subsequence_pair<ORDER>(cur, T, lut_g, walk, begin, unit, cp_tab, span32, row, &cp_state[0][slot], &cp_count[0][slot], true)
				                      : walk_subsequence<ORDER>(cur, lut_sa, lut_g, walk, begin, unit, cp_tab, span32, row, &cp_state[0][slot], &cp_count[0][slot], true);
				if(hit) active = false;
			}
			if(!__syncthreads_or(active ? 1 : 0)) break;
		}
		if(my_sub >= 0 && my_sub < end_sub) {
			uint32_t e = cp_state[kCp - 1][tid];
			if(PAIR && ORDER) e = (e & 0xff00u) | lds_u8(T.live + (e & 255u));   // checkpoints hold the row: report the context byte
			if(tid >= warm) {
				uint32_t total = 0;
#pragma unroll
				for(int j = 0; j < kCp; ++j) total += cp_count[j][tid];
				state[my_sub] = e;
				count[my_sub] = total;
			} else if(tid == warm - 1) {
				seam[chunk] = e;   // boundary state as this chunk saw it
			}
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------------------
// D2: chunk seams. flags[slot] is raised when some chunk's END state changed (the next pass must re-check).
// ---------------------------------------------------------------------------------------------------------
template <int ORDER>
__global__ void dec_seam_kernel(const uint32_t* __restrict__ words, uint64_t n_bits, uint64_t buf_bytes, const uint16_t* __restrict__ lut_g,
                                const uint32_t* __restrict__ walk, uint32_t* state, uint32_t* count, uint32_t* seam,
                                uint32_t sub_bits, uint64_t n_subs, uint32_t n_chunks, uint32_t chunk_subs, uint32_t phase, uint32_t* flag) {
	__shared__ __align__(16) uint8_t seam_ring[128 * kRingBytesPerThread];   // launched with 128 threads
	const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x + 1;
	if(chunk >= n_chunks) return;
	const uint64_t first = uint64_t(chunk) * chunk_subs;
	const uint32_t recorded = *(volatile uint32_t*) (state + first - 1);
	if(seam[chunk] == recorded) return;
	seam[chunk] = recorded;
	const uint64_t last = first + chunk_subs < n_subs ? first + chunk_subs : n_subs;
	const uint64_t origin = first * sub_bits + phase;
	Cursor cur;
	cur.words = words;
	cur.n_bytes = buf_bytes;
	cur.attach(uint32_t(__cvta_generic_to_shared(seam_ring)));
	uint32_t pos = recorded >> 8, ctx = recorded & 255u;
	cur.seek(origin + pos, pos);
	bool merged = false;
	for(uint64_t k = first; k < last; ++k) {
		uint64_t e = (k + 1) * sub_bits + phase;
		if(e > n_bits) e = n_bits;
		const uint32_t lim = uint32_t(e - origin);
		uint32_t cnt = 0;
		decode_until<ORDER, false>(cur, 0u, lut_g, walk, lim, ctx, cnt);
		pos = cur.pos;
		const uint32_t st = pack_cp(pos - lim, ORDER ? ctx : 0u);
		count[k] = cnt;
		if(state[k] == st) { merged = true; break; }
		state[k] = st;
	}
	if(!merged) *flag = 1;
}

// ---------------------------------------------------------------------------------------------------------
// D3: per-chunk totals and their exclusive scan
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_total_kernel(const uint32_t* __restrict__ count, uint64_t n_subs, uint32_t chunk_subs,
                                                         uint32_t skip_subs, unsigned long long* __restrict__ chunk_total,
                                                         uint32_t* __restrict__ prefix) {
	// one CTA per chunk: the chunk's symbol total, and for every subsequence the symbols of the chunk before it
	// (D4 adds the chunk's base and has its output offset without any scan of its own)
	__shared__ uint32_t part[8];
	__shared__ uint32_t carry_s;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t first = uint64_t(blockIdx.x) * chunk_subs;
	const uint64_t last = first + chunk_subs < n_subs ? first + chunk_subs : n_subs;
	if(threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	for(uint64_t base = first; base < last; base += 256) {
		const uint64_t k = base + threadIdx.x;
		const uint32_t c = (k < last && k >= skip_subs) ? count[k] : 0u;   // a shard's leading warm-up subsequences belong to its predecessor
		uint32_t incl = c;
		for(int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
			if(lane >= uint32_t(d)) incl += t;
		}
		if(lane == 31) part[warp] = incl;
		__syncthreads();
		uint32_t before = carry_s;
		for(uint32_t w = 0; w < warp; ++w) before += part[w];
		if(k < last) prefix[k] = before + incl - c;
		__syncthreads();
		if(threadIdx.x == 255) carry_s = before + incl;
		__syncthreads();
	}
	if(threadIdx.x == 0) chunk_total[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(1024) dec_scan_kernel(const unsigned long long* __restrict__ chunk_total,
                                                         unsigned long long* __restrict__ chunk_base, uint32_t n_chunks,
                                                         uint64_t out_capacity, const uint32_t* __restrict__ flags, int last_flag,
                                                         const uint32_t* __restrict__ state, uint64_t n_subs, uint32_t skip_subs,
                                                         unsigned long long* __restrict__ result) {
	__shared__ unsigned long long warp_tot[32];
	__shared__ unsigned long long carry_s;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if(tid == 0) carry_s = 0;
	__syncthreads();
	for(uint32_t base = 0; base < n_chunks; base += 1024) {
		const uint32_t i = base + tid;
		const unsigned long long v = i < n_chunks ? chunk_total[i] : 0ull;
		unsigned long long incl = v;
		for(int d = 1; d < 32; d <<= 1) {
			const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
			if(lane >= uint32_t(d)) incl += t;
		}
		if(lane == 31) warp_tot[warp] = incl;
		__syncthreads();
		unsigned long long before = carry_s;
		for(uint32_t w = 0; w < warp; ++w) before += warp_tot[w];
		if(i < n_chunks) chunk_base[i] = before + incl - v;
		__syncthreads();
		if(tid == 1023) carry_s = before + incl;
		__syncthreads();
	}
	if(tid == 0) {
		const unsigned long long total = carry_s;
		chunk_base[n_chunks] = total;
		result[0] = total;
		long long status = 0;
		if(last_flag >= 0 && flags[last_flag]) status = MH_ERR_NOT_CONVERGED;
		else if(total > out_capacity) status = MH_ERR_CAPACITY;
		result[1] = (unsigned long long) status;
		// shard seams: [63:32] the state at the ownership start as the warm-up saw it, [31:0] the state at the end
		const unsigned long long view = skip_subs ? state[skip_subs - 1] : 0u;
		result[3] = (view << 32) | state[n_subs - 1];
	}
}

// ---------------------------------------------------------------------------------------------------------
// D4: final decode + write
// ---------------------------------------------------------------------------------------------------------
template <int ORDER, bool PAIR>
__global__ void __launch_bounds__(PAIR ? kDecWriteMaxThreads : kDecThreads, 1) dec_write_kernel(   // launched with decode_write_threads()
    const uint32_t* __restrict__ words, uint64_t n_bits, uint64_t buf_bytes, uint32_t start0, const uint16_t* __restrict__ lut_g,
    const uint32_t* __restrict__ walk, const uint32_t* __restrict__ pair_g, uint32_t pair_rows, uint32_t pair_ctx_rows,
    const uint32_t* __restrict__ state, const uint32_t* __restrict__ count, const uint32_t* __restrict__ prefix,
    const unsigned long long* __restrict__ chunk_base, uint8_t* __restrict__ out, uint32_t sub_bits, uint64_t n_subs, uint32_t chunk_subs,
    uint32_t skip_subs, uint32_t stream_end, unsigned long long* result, unsigned long long* ticket, uint32_t lut_smem_bytes) {
	extern __shared__ __align__(16) uint16_t lut_s[];   // table, then the payload rings, then the output rings
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if(result[1] != 0) return;   // capacity / convergence error decided by D3: write nothing
	PairTab T = {};
	if(PAIR) {
		T = stage_pair_table(reinterpret_cast<uint32_t*>(lut_s), pair_g, pair_rows, pair_ctx_rows);
	} else {
		const uint32_t n16 = ORDER ? 65536u : 256u;
		const uint4* src = reinterpret_cast<const uint4*>(lut_g);
		uint4* dst = reinterpret_cast<uint4*>(lut_s);
		for(uint32_t i = tid; i < n16 / 8; i += blockDim.x) dst[i] = src[i];
	}
	__syncthreads();
	uint32_t lut_sa;
	asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(lut_sa) : "l"(lut_s));
	Cursor cur;
	cur.words = words;
	cur.n_bytes = buf_bytes;
	cur.attach(lut_sa + lut_smem_bytes);
	OutStage os;
	os.ring_warp = lut_sa + lut_smem_bytes + blockDim.x * kRingBytesPerThread + warp * 32u * kOutRingBytes;
	os.ring = os.ring_warp + lane * kOutRingBytes;
	bool clean = true;
	// Work is handed out per warp, 32 consecutive subsequences at a time, from an atomic counter: no barrier in
	// the loop, no tail wait for the slowest warp of a CTA. A subsequence's output offset is its chunk's base (D3's
	// scan) plus the symbols of the chunk before it (D3's prefix).
	for(;;) {
		unsigned long long base = 0;
		if(lane == 0) base = atomicAdd(ticket, 32ull);
		base = __shfl_sync(0xffffffffu, base, 0);
		if(base >= n_subs) break;
		const uint64_t k = base + lane;
		const bool mine = k < n_subs && k >= skip_subs;
		const uint32_t c = mine ? count[k] : 0u;
		// every lane goes through the emitter (the pair path writes out warp-wide); lanes without symbols pass count 0
		const uint32_t start = !mine ? 0u : (k == 0 ? (start0 & 255u) : state[k - 1]);
		const uint64_t origin = k * sub_bits + (start0 >> 8);
		uint64_t e = origin + sub_bits;
		if(e > n_bits) e = n_bits;
		uint32_t pos = start >> 8;
		const uint32_t ctx = start & 255u;
		const uint32_t lim = uint32_t(e - origin);
		uint8_t* dst = out;
		if(c) {
			cur.set_reach(origin + sub_bits + 64);   // the last codeword may run a few bits past the subsequence
			cur.seek(origin + pos, pos);
			dst = out + (chunk_base[k / chunk_subs] + prefix[k]);
		}
		if(PAIR) {
			clean &= decode_emit_pair<ORDER>(cur, T, lut_g, walk, c, ctx, dst, stream_end && k == n_subs - 1, os);
		} else if(c) {
			clean &= decode_emit<ORDER>(cur, lut_sa, lut_g, walk, c, ctx, dst);
		}
		if(mine) {
			if(c) pos = cur.pos;
			// D1 counted the symbols that start before `lim`: decoding that many must land on or after it, and on
			// the very end of the payload for the last subsequence (else the last codeword runs past the stream)
			if(pos < lim) clean = false;
			if(stream_end && k == n_subs - 1 && origin + pos != n_bits) clean = false;
		}
	}
	if(!clean) result[2] = (unsigned long long) (long long) MH_ERR_CORRUPT_STREAM;
}

}  // namespace

static uint32_t decode_write_threads() {
	const long long t = tunable(kTunDecWriteThreads);   // experiments
	const int v = t > 0 ? int(t) : kDecWriteMaxThreads;
	return (v >= 64 && v <= kDecWriteMaxThreads && v % 32 == 0) ? uint32_t(v) : uint32_t(kDecWriteMaxThreads);
}

uint32_t decode_sub_bits(int order, uint64_t n_bits) {
	// Subsequence size. Longer subsequences amortise the self-synchronisation overlap (Markov streams re-synchronise
	// ~10x slower than plain Huffman streams) but there must be enough of them to fill the machine:
	// the largest candidate that still yields kDecTargetSubs subsequences, else the smallest.
	// The tunable override exists for experiments and tests.
	const long long v = tunable(order ? kTunDecSubBitsMarkov : kTunDecSubBitsHuffman);
	if(v >= kDecMinSubBits && v % 256 == 0 && v <= (1 << 16)) return uint32_t(v);
	const uint32_t largest = order ? kDecMaxSubBitsMarkov : kDecMaxSubBitsHuffman;
	const uint32_t smallest = order ? 1024u : 512u;
	uint32_t sub = largest;
	while(sub > smallest && n_bits / sub < kDecTargetSubs) sub >>= 1;
	return sub;
}

uint64_t decode_max_subs(uint64_t max_payload_bytes) {
	// workspace bound: a candidate below the largest is only chosen while it yields fewer than 2 x kDecTargetSubs
	const uint64_t bits = max_payload_bytes * 8 + 64;
	uint64_t cap = 2 * kDecTargetSubs + 2;
	for(int order = 0; order < 2; ++order) {
		const uint64_t by_largest = bits / decode_sub_bits(order, ~0ull >> 8) + 2;   // also honours an env override
		if(by_largest > cap) cap = by_largest;
	}
	return cap;
}

namespace {

template <int ORDER, bool PAIR>
int run_decode(const uint32_t* words, uint64_t n_bits, uint64_t buf_bytes, uint32_t start0, uint32_t skip_subs, uint32_t stream_end,
               const mh_dectable* dt, uint8_t* d_out, uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws,
               cudaStream_t st, int fix_iters, uint32_t sub_bits) {
	const uint64_t n_subs = (n_bits - (start0 >> 8) + sub_bits - 1) / sub_bits;   // the grid starts at the stream's first bit
	// the next chunk re-decodes the last `warm` subsequences (>= 8192 bits) of this one as warm-up
	uint32_t warm = 8192 / sub_bits;
	warm = warm < 1 ? 1 : (warm > uint32_t(kDecWarmSubs) ? uint32_t(kDecWarmSubs) : warm);
	const int sms = sm_count();
	// D1 threads per CTA = subsequences per chunk + warm-up: kDecThreads, or fewer when that would leave SMs without a chunk
	// (the 256 MiB Fibonacci stream has 86 k subsequences: 85 chunks of 1016, or 148 of 580)
	uint32_t d1_threads = uint32_t(kDecThreads);
	if((n_subs + (kDecThreads - warm) - 1) / (kDecThreads - warm) < uint64_t(sms)) {
		const uint64_t per = ((n_subs + sms - 1) / sms + warm + 31) & ~uint64_t(31);
		d1_threads = uint32_t(per < 256 ? 256 : (per > uint64_t(kDecThreads) ? uint64_t(kDecThreads) : per));
	}
	const uint32_t chunk_subs = d1_threads - warm;
	const uint64_t chunks64 = (n_subs + chunk_subs - 1) / chunk_subs;
	if(n_subs > ws->dec_subs_cap || chunks64 > ws->dec_chunks_cap) return MH_ERR_WORKSPACE;
	const uint32_t n_chunks = uint32_t(chunks64);
	const size_t lut_bytes = PAIR ? size_t(pair_table_bytes(dt->pair_rows, dt->pair_ctx_rows)) : (ORDER ? 65536 * 2 : 256 * 2);
	static std::atomic<uint64_t> attr_done{0};   // one per template instantiation, one bit per device
	if(first_use_on_device(attr_done)) {
		const int ring_bytes = kDecThreads * int(kRingBytesPerThread);
		MH_CUDA(cudaFuncSetAttribute(dec_sync_kernel<ORDER, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (PAIR ? kDecPairBytes : int(lut_bytes)) + ring_bytes));
		MH_CUDA(cudaFuncSetAttribute(dec_write_kernel<ORDER, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                             PAIR ? max_smem_optin() - 1024 : int(lut_bytes) + ring_bytes));   // pair path: the launcher fits the thread count
	}
	// Checkpoints inside a subsequence, in 1/32 of its length. A walker that took over its successor's subsequence stops
	// at the first checkpoint where its state equals the recorded one: dense near the start, where re-synchronisation
	// usually happens (1, 2, 4, 8, 12, 16, 24, 32: -8 % on D1 with -h, where streams re-synchronise within tens of bits;
	// no change for Markov text), or evenly spaced (4, 8, .. 32) with the tunable dec_cp_geo = 0.
	const uint64_t cp_tab = tunable(kTunDecCpGeo) != 0 ? 0x2018100c08040201ull : 0x201c1814100c0804ull;
	const uint32_t grid = n_chunks < uint32_t(sms) ? n_chunks : uint32_t(sms);
	MH_CUDA(cudaMemsetAsync(ws->dec_flags, 0, 8 * sizeof(uint32_t), st));
	MH_CUDA(cudaMemsetAsync(ws->counters + 4, 0, sizeof(unsigned long long), st));   // D4's work ticket
	{
		ProfScope p("dec_sync_kernel", st);
		dec_sync_kernel<ORDER, PAIR><<<grid, d1_threads, lut_bytes + size_t(d1_threads) * kRingBytesPerThread, st>>>(words, n_bits, buf_bytes, start0, dt->d_lut, dt->d_walk, dt->d_pair,
		    dt->pair_rows, dt->pair_ctx_rows, ws->dec_state, ws->dec_count, ws->dec_seam, sub_bits, n_subs, n_chunks, warm, uint32_t(lut_bytes), cp_tab);
	}
	count_launch(1);
	int last_flag = -1;
	if(n_chunks > 1) {
		if(fix_iters < 1) fix_iters = 1;
		for(int it = 0; it < fix_iters; ++it) {
			const int slot = it & 7;   // the flag slots are reused round-robin
			if(it >= 8) MH_CUDA(cudaMemsetAsync(ws->dec_flags + slot, 0, sizeof(uint32_t), st));
			ProfScope p("dec_seam_kernel", st);
			dec_seam_kernel<ORDER><<<(n_chunks - 1 + 127) / 128, 128, 0, st>>>(words, n_bits, buf_bytes, dt->d_lut, dt->d_walk, ws->dec_state,
			    ws->dec_count, ws->dec_seam, sub_bits, n_subs, n_chunks, chunk_subs, start0 >> 8, ws->dec_flags + slot);
			count_launch(1);
			last_flag = slot;
		}
	}
	{
		ProfScope p("dec_total_kernel", st);
		dec_total_kernel<<<n_chunks, 256, 0, st>>>(ws->dec_count, n_subs, chunk_subs, skip_subs, (unsigned long long*) ws->dec_chunk_total, ws->dec_prefix);
	}
	{
		ProfScope p("dec_scan_kernel", st);
		dec_scan_kernel<<<1, 1024, 0, st>>>((const unsigned long long*) ws->dec_chunk_total, (unsigned long long*) ws->dec_chunk_base,
		    n_chunks, out_capacity, ws->dec_flags, last_flag, ws->dec_state, n_subs, skip_subs, d_result);
	}
	{
		ProfScope p("dec_write_kernel", st);
		// threads per CTA: as many warps as the table and the per-thread rings leave room for, one CTA per SM
		const size_t per_thread = kRingBytesPerThread + (PAIR ? kOutRingBytes : 0u);
		uint32_t wt = decode_write_threads();
		while(wt > 128 && lut_bytes + size_t(wt) * per_thread + 1024 > size_t(max_smem_optin())) wt -= 32;
		const uint64_t warps_needed = (n_subs + 31) / 32;
		// a stream with few subsequences: smaller CTAs on every SM rather than full ones on some of them (the 256 MiB
		// Fibonacci stream has 86 k subsequences: 84 CTAs of 1024 threads, or 148 of 608)
		if(tunable(kTunDecWriteThreads) <= 0 && warps_needed < uint64_t(sms) * (wt / 32)) {
			const uint32_t per_sm = uint32_t((warps_needed + sms - 1) / sms) * 32u;
			wt = per_sm < 128u ? 128u : (per_sm < wt ? per_sm : wt);
		}
		uint64_t wgrid = (warps_needed + wt / 32 - 1) / (wt / 32);
		if(wgrid > uint64_t(sms)) wgrid = uint64_t(sms);
		dec_write_kernel<ORDER, PAIR><<<unsigned(wgrid), wt, lut_bytes + size_t(wt) * per_thread, st>>>(words, n_bits, buf_bytes, start0, dt->d_lut, dt->d_walk,
		    dt->d_pair, dt->pair_rows, dt->pair_ctx_rows, ws->dec_state, ws->dec_count, ws->dec_prefix, (const unsigned long long*) ws->dec_chunk_base, d_out, sub_bits,
		    n_subs, chunk_subs, skip_subs, stream_end, d_result, reinterpret_cast<unsigned long long*>(ws->counters + 4), uint32_t(lut_bytes));
	}
	count_launch(3);
	MH_CUDA(cudaGetLastError());
	return MH_OK;
}

// The pair table is used whenever it exists (<= 63 live contexts); the tunable dec_pair = 0 forces the u16 LUT (tests, experiments).
int dispatch_decode(const uint32_t* words, uint64_t n_bits, uint64_t buf_bytes, uint32_t start0, uint32_t skip_subs, uint32_t stream_end,
                    const mh_dectable* dt, uint8_t* d_out, uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws,
                    cudaStream_t st, int fix_iters, uint32_t sub_bits) {
	const bool pair = dt->pair_rows != 0 && tunable(kTunDecPair) != 0;
	if(dt->order) {
		if(pair) return run_decode<1, true>(words, n_bits, buf_bytes, start0, skip_subs, stream_end, dt, d_out, out_capacity, d_result, ws, st, fix_iters, sub_bits);
		return run_decode<1, false>(words, n_bits, buf_bytes, start0, skip_subs, stream_end, dt, d_out, out_capacity, d_result, ws, st, fix_iters, sub_bits);
	}
	if(pair) return run_decode<0, true>(words, n_bits, buf_bytes, start0, skip_subs, stream_end, dt, d_out, out_capacity, d_result, ws, st, fix_iters, sub_bits);
	return run_decode<0, false>(words, n_bits, buf_bytes, start0, skip_subs, stream_end, dt, d_out, out_capacity, d_result, ws, st, fix_iters, sub_bits);
}

}  // namespace

int launch_decode(const uint8_t* d_bits, uint64_t bit_base, uint64_t n_bits, uint8_t prev0, const mh_dectable* dt, uint8_t* d_out,
                  uint64_t out_capacity, unsigned long long* d_result, mh_workspace* ws, cudaStream_t st, int fix_iters) {
	if(!dt || !dt->d_lut || !d_result || (!d_bits && n_bits)) return MH_ERR_INVALID_ARG;
	if(reinterpret_cast<uint64_t>(d_bits) & 3) return MH_ERR_INVALID_ARG;
	if(!ws || !ws->dec_state) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(d_result, 0, 4 * sizeof(unsigned long long), st));
	if(n_bits == 0) return MH_OK;
	const uint32_t* words = reinterpret_cast<const uint32_t*>(d_bits);
	// the payload starts at bit (bit_base & 7) of d_bits[0]: the kernels see a stream that ends at bit0 + n_bits and
	// whose first subsequence starts, exactly known, at (bit0, prev0)
	const uint32_t bit0 = uint32_t(bit_base & 7);
	const uint32_t start0 = (bit0 << 8) | prev0;
	const uint64_t end_bit = n_bits + bit0;
	const uint64_t buf_bytes = (end_bit + 7) >> 3;
	const uint32_t sub_bits = decode_sub_bits(dt->order, n_bits);
	return dispatch_decode(words, end_bit, buf_bytes, start0, 0, 1, dt, d_out, out_capacity, d_result, ws, st, fix_iters, sub_bits);
}

int launch_decode_shard(const uint8_t* d_bits, uint32_t start_bit, uint64_t n_bits, uint64_t buf_bytes, int exact_start, uint8_t prev0,
                        uint32_t warm_bits, int stream_end, const mh_dectable* dt, uint8_t* d_out, uint64_t out_capacity,
                        unsigned long long* d_result, mh_workspace* ws, cudaStream_t st, int fix_iters) {
	if(!dt || !dt->d_lut || !d_result || !d_bits || start_bit > 31) return MH_ERR_INVALID_ARG;
	if(reinterpret_cast<uint64_t>(d_bits) & 3) return MH_ERR_INVALID_ARG;
	if(!ws || !ws->dec_state) return MH_ERR_WORKSPACE;
	MH_CUDA(cudaMemsetAsync(d_result, 0, 4 * sizeof(unsigned long long), st));
	if(n_bits == 0) return MH_OK;
	const uint64_t end_bit = n_bits + start_bit;
	if(buf_bytes * 8 < end_bit) return MH_ERR_INVALID_ARG;
	if(warm_bits >= n_bits) return MH_ERR_INVALID_ARG;
	const uint32_t sub_bits = decode_sub_bits(dt->order, n_bits - warm_bits);   // sized by the bits the shard owns
	if(warm_bits % sub_bits) return MH_ERR_INVALID_ARG;                         // warm-up = whole subsequences
	const uint32_t skip_subs = warm_bits / sub_bits;
	const uint32_t* words = reinterpret_cast<const uint32_t*>(d_bits);
	// a shard that does not know its start state guesses the context; its leading skip_subs subsequences are warm-up
	const uint32_t start0 = (start_bit << 8) | (exact_start ? uint32_t(prev0) : uint32_t(' '));
	return dispatch_decode(words, end_bit, buf_bytes, start0, skip_subs, stream_end ? 1 : 0, dt, d_out, out_capacity, d_result, ws, st, fix_iters, sub_bits);
}

}  // namespace mh
