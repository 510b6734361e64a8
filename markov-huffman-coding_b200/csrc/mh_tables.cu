// mh_tables.cu — the encoder's tables built ON THE DEVICE from the device-resident histogram.
//
// The reference builds its (up to) 256 Huffman trees on the host (huffman_table::build, src/huffman.cpp:131-164, on
// min_pq, src/min_pq.tpp:4-52) — microsecond work, but between the histogram and the encoder it is a device-to-host
// copy of the counts, ~0.3 ms of host work and a host-to-device copy of the tables, all on the critical path of a
// compress. Here one warp per context runs the SAME algorithm next to the data: the reference's array heap with its
// strict comparisons (the tie-breaking is part of the format), int32 wrapping weights, the
// "shallower subtree goes left" swap (src/huffman.cpp:147-149), the single-symbol fake root (:154-162) and the
// pre-order code assignment (:97-123) — and writes the encoder's two flat tables straight into the codebook handle:
//   wide   u64[ntab * 256]  len << 56 | code                                  (CodingTable::flatten_codebook)
//   ctx    u32[rows * 256]  len << 27 | next row << 16 | code, + the null row  (CodingTable::flatten_ctx)
// The host still builds its own mh_table from the same counts (for the table file and the decoder's tables), but off
// the critical path, while the encoder runs; tests compare the device-built tables with the host-built ones bit by bit.
// The serial part of a tree (heap operations) is done by lane 0 of the context's warp in shared memory; the warp
// reads the counts and writes the table rows together. All contexts run in parallel on different SMs.
//
// meta (u32[8], device): [0] context rows incl. the null row, [1] status (0, or a negative mh_status: a live count that
// wrapped to 0, a codeword longer than 56 bits), [2] longest codeword, [3] live contexts, [4..5] u64 sum of count x codeword
// length (the payload's bits, up to wrapped counts), [6..7] u64 sum of the counts.
#include "mh_internal.hpp"

namespace mh {
namespace {

struct TreeScratch {
	unsigned long long heap[256];   // the array heap: weight << 32 | node; after the heap phase: the paths of the internal nodes
	int16_t left[512], right[512];
	int32_t weight[512];
	int16_t height[512];
	uint8_t symbol[256];      // of leaf i (leaves are nodes 0 .. n_leaves - 1, in ascending symbol order)
	uint8_t len[256];
	unsigned long long code[256];
};

// The heap's entries are one 64-bit word each — weight (int32, the reference's `int`) in the high half, node in the low
// half — and an element that moves several levels travels through a hole (one store per level and one at the end)
// instead of being swapped level by level: the same comparisons in the same order, so the same arrangement as the
// reference's swaps, at half the shared-memory round trips of the one lane that runs this.
__device__ __forceinline__ int32_t heap_weight(unsigned long long e) { return int32_t(uint32_t(e >> 32)); }

// min_pq::insert (src/min_pq.tpp:4-8, swim :29-36): move up while the parent is strictly heavier
__device__ __forceinline__ void heap_push(TreeScratch& S, int& hs, int32_t w, int node) {
	int i = hs++;
	const unsigned long long me = (static_cast<unsigned long long>(uint32_t(w)) << 32) | uint32_t(node);
	while(i > 0) {
		const int up = (i - 1) >> 1;
		const unsigned long long u = S.heap[up];
		if(!(heap_weight(u) > w)) break;
		S.heap[i] = u;
		i = up;
	}
	S.heap[i] = me;
}

// min_pq::pop_min (src/min_pq.tpp:9-27, sink :38-52): the right child only if strictly lighter than the left one, and a
// swap only if the candidate is strictly lighter than the node
__device__ __forceinline__ int heap_pop(TreeScratch& S, int& hs) {
	const int top = int(uint32_t(S.heap[0]));
	--hs;
	const unsigned long long me = S.heap[hs];
	const int32_t w = heap_weight(me);
	int i = 0;
	for(;;) {
		const int l = 2 * i + 1, r = l + 1;
		if(l >= hs) break;
		unsigned long long pick_e = S.heap[l];
		int pick = l;
		if(r < hs) {
			const unsigned long long right_e = S.heap[r];
			if(heap_weight(right_e) < heap_weight(pick_e)) { pick = r; pick_e = right_e; }
		}
		if(!(heap_weight(pick_e) < w)) break;
		S.heap[i] = pick_e;
		i = pick;
	}
	S.heap[i] = me;
	return top;
}

// One tree from S.weight[0 .. n_leaves) / S.symbol (lane 0). Fills S.len / S.code; returns 0 or a negative mh_status.
__device__ int build_tree(TreeScratch& S, int n_leaves, uint32_t& max_len) {
	int hs = 0, n = n_leaves;
	for(int i = 0; i < n_leaves; ++i) {
		S.left[i] = S.right[i] = -1;
		S.height[i] = 0;
		heap_push(S, hs, S.weight[i], i);   // ascending symbol order fixes the initial heap layout (src/huffman.cpp:134-138)
	}
	while(hs > 1) {
		int a = heap_pop(S, hs), b = heap_pop(S, hs);
		if(S.height[a] > S.height[b]) { const int t = a; a = b; b = t; }   // shallower subtree on the left (src/huffman.cpp:147-149)
		S.left[n] = int16_t(a);
		S.right[n] = int16_t(b);
		S.weight[n] = int32_t(uint32_t(S.weight[a]) + uint32_t(S.weight[b]));   // int weight: wraps (src/tree.h:14,20)
		S.height[n] = int16_t((S.height[a] > S.height[b] ? S.height[a] : S.height[b]) + 1);
		heap_push(S, hs, S.weight[n], n);
		++n;
	}
	const int root = heap_pop(S, hs);
	if(S.left[root] < 0) {   // one live symbol: fake root over two copies of the leaf, the right copy ("1") wins (src/huffman.cpp:154-162)
		S.len[S.symbol[root]] = 1;
		S.code[S.symbol[root]] = 1ull;
		max_len = 1;
		return 0;
	}
	// Codes: left edge 0, right edge 1 — the reference's pre-order walk (src/huffman.cpp:97-123) gives every leaf the path
	// from the root, and the order the nodes are visited in does not change a path. A node's children were made before
	// it, so their indices are smaller: ONE pass from the root down over the internal nodes hands every child its depth
	// and path — no stack, a fifth of the walk's shared-memory round trips (the tree build is serial work of one lane).
	// The heap phase is over, so its arrays carry the pass: height[node] <- depth, heap[node - n_leaves] <- path (internal
	// nodes only: fewer than 256).
	int status = 0;
	uint32_t longest = 0;
	S.height[root] = 0;
	S.heap[root - n_leaves] = 0;
	for(int node = root; node >= n_leaves; --node) {
		const uint32_t depth = uint32_t(S.height[node]) + 1u;
		const unsigned long long path = S.heap[node - n_leaves];
#pragma unroll
		for(int side = 0; side < 2; ++side) {
			const int child = side ? S.right[node] : S.left[node];
			const unsigned long long cp = (path << 1) | static_cast<unsigned long long>(side);
			if(child < n_leaves) {
				S.len[S.symbol[child]] = uint8_t(depth);
				S.code[S.symbol[child]] = cp;
				if(depth > longest) longest = depth;
				if(depth > uint32_t(MH_MAX_CODE_BITS)) status = MH_ERR_CODE_TOO_LONG;
			} else {
				S.height[child] = int16_t(depth);
				S.heap[child - n_leaves] = cp;
			}
		}
	}
	max_len = longest;
	return status;
}

// Which contexts have a tree, and the wrap check: one warp per context row, coalesced. flags[p]: bit 0 live, bit 1 a
// live count that the reference's int counter shows as 0.
__global__ void __launch_bounds__(32) tables_scan_kernel(const unsigned long long* __restrict__ counts, uint8_t* __restrict__ flags) {
	const uint32_t p = blockIdx.x, lane = threadIdx.x;
	const unsigned long long* row = counts + size_t(p) * 256;
	bool live = false, wrapped = false;
#pragma unroll
	for(int k = 0; k < 8; ++k) {
		const unsigned long long c = row[k * 32 + lane];
		live |= uint32_t(c) != 0;                        // the reference counts in `int` (src/main.cpp:166,174): the low 32 bits
		wrapped |= c != 0 && uint32_t(c) == 0;           // the reference would lose this symbol
	}
	live = __any_sync(0xffffffffu, live);
	wrapped = __any_sync(0xffffffffu, wrapped);
	if(lane == 0) flags[p] = uint8_t((live ? 1 : 0) | (wrapped ? 2 : 0));
}

// One warp per context: tree, codes, and the context's rows of the two encoder tables.
__global__ void __launch_bounds__(32) tables_build_kernel(const unsigned long long* __restrict__ counts, int order, uint32_t* __restrict__ meta,
                                                           const uint8_t* __restrict__ flags, unsigned long long* __restrict__ enc, uint32_t* __restrict__ ctx,
                                                           uint32_t max_ctx_rows) {
	__shared__ TreeScratch S;
	__shared__ uint8_t rank[256];   // row of every byte value as a context (meaningful where has[] is set and the rows fit a byte)
	__shared__ uint8_t has[256];    // the context has a tree
	__shared__ int s_status;
	const uint32_t p = blockIdx.x, lane = threadIdx.x;
	// every block ranks the live contexts itself (256 flag bytes): lane l owns the values 8 l .. 8 l + 7
	uint32_t mine = 0, bad = 0;
	for(uint32_t k = 0; k < 8; ++k) {
		const uint32_t f = (order || lane * 8 + k == 0) ? flags[lane * 8 + k] : 0u;
		mine |= (f & 1u) << k;
		bad |= f & 2u;
	}
	uint32_t before = __popc(mine);
	for(int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, before, d);
		if(lane >= uint32_t(d)) before += t;
	}
	const uint32_t live = __shfl_sync(0xffffffffu, before, 31), rows = live + 1;
	before -= __popc(mine);
	for(uint32_t k = 0; k < 8; ++k) {
		rank[lane * 8 + k] = uint8_t(before);
		has[lane * 8 + k] = uint8_t((mine >> k) & 1u);
		before += (mine >> k) & 1u;
	}
	const bool wrapped = __any_sync(0xffffffffu, bad != 0);
	__syncwarp();
	const bool ctx_fits = rows <= max_ctx_rows;
	const uint32_t my_row = rank[p & 255u];
	const bool my_tree = has[p & 255u] != 0;
	auto next_of = [&](uint32_t c) -> uint32_t { return order ? (has[c] ? uint32_t(rank[c]) : live) : 0u; };
	if(blockIdx.x == gridDim.x - 1) {   // the extra block writes the null row and the summary
		if(ctx_fits)
			for(uint32_t c = lane; c < 256; c += 32) ctx[size_t(live) * 256 + c] = next_of(c) << 16;
		if(lane == 0) {
			meta[0] = rows;
			meta[3] = live;
			if(wrapped) meta[1] = uint32_t(MH_ERR_COUNT_WRAPPED);
		}
		return;
	}
	unsigned long long* enc_row = enc + size_t(p) * 256;
	if(!my_tree || wrapped) {   // no tree for this context (or the counts are unusable): no codewords
		for(uint32_t c = lane; c < 256; c += 32) enc_row[c] = 0;
		return;
	}
	const unsigned long long* row = counts + size_t(p) * 256;
	for(uint32_t c = lane; c < 256; c += 32) { S.len[c] = 0; S.code[c] = 0; }
	__syncwarp();
	// the leaves in ascending symbol order (src/huffman.cpp:134-138), gathered by the whole warp: eight coalesced loads per lane
	// instead of 256 loads by lane 0
	int n = 0;
	for(int k = 0; k < 8; ++k) {
		const int sym = k * 32 + int(lane);
		const int32_t w = int32_t(uint32_t(row[sym]));
		const uint32_t m = __ballot_sync(0xffffffffu, w != 0);   // `if(counts[i])`: a wrapped-negative count is still a leaf
		if(w != 0) {
			const int at = n + __popc(m & ((1u << lane) - 1u));
			S.weight[at] = w;
			S.symbol[at] = uint8_t(sym);
		}
		n += __popc(m);
	}
	__syncwarp();
	if(lane == 0) {
		uint32_t longest = 0;
		s_status = build_tree(S, n, longest);
		atomicMax(meta + 2, longest);
		if(s_status) meta[1] = uint32_t(s_status);
	}
	__syncwarp();
	unsigned long long bits = 0, symbols = 0;   // this context's share of the payload: the encoder sizes itself by the mean codeword
	for(uint32_t c = lane; c < 256; c += 32) {
		const uint32_t len = S.len[c];
		const unsigned long long code = S.code[c];
		enc_row[c] = len ? (static_cast<unsigned long long>(len) << 56) | code : 0ull;
		if(ctx_fits)
			ctx[size_t(my_row) * 256 + c] = len > 16 ? (31u << 27) | (next_of(c) << 16) : (len << 27) | (next_of(c) << 16) | (len ? uint32_t(code) : 0u);
		bits += row[c] * len;
		symbols += row[c];
	}
	for(int d = 16; d; d >>= 1) {
		bits += __shfl_xor_sync(0xffffffffu, bits, d);
		symbols += __shfl_xor_sync(0xffffffffu, symbols, d);
	}
	if(lane == 0) {
		atomicAdd(reinterpret_cast<unsigned long long*>(meta + 4), bits);
		atomicAdd(reinterpret_cast<unsigned long long*>(meta + 6), symbols);
	}
}

}  // namespace

// Builds cb's device tables from d_counts (u64[256] order 0 / u64[65536] order 1, device) on `st`. The handle must own
// its device buffers (mh_codebook_create / a session / a comm). Nothing is copied to the host and the host does not wait.
int launch_build_codebook(const unsigned long long* d_counts, int order, mh_codebook* cb, cudaStream_t st) {
	if(!d_counts || !cb || (order != 0 && order != 1)) return MH_ERR_INVALID_ARG;
	if(!cb->d_enc) MH_CUDA(cudaMalloc(&cb->d_enc, 65536 * sizeof(uint64_t)));
	if(!cb->d_ctx) MH_CUDA(cudaMalloc(&cb->d_ctx, size_t(kEncCtxMaxRows) * 256 * sizeof(uint32_t)));
	if(!cb->d_meta) MH_CUDA(cudaMalloc(&cb->d_meta, 8 * sizeof(uint32_t) + 256));
	uint8_t* d_flags = reinterpret_cast<uint8_t*>(cb->d_meta + 8);
	MH_CUDA(cudaMemsetAsync(cb->d_meta, 0, 8 * sizeof(uint32_t), st));
	{
		ProfScope p("tables_scan_kernel", st);
		tables_scan_kernel<<<order ? 256 : 1, 32, 0, st>>>(d_counts, d_flags);
	}
	{
		ProfScope p("tables_build_kernel", st);
		tables_build_kernel<<<order ? 257 : 2, 32, 0, st>>>(d_counts, order, cb->d_meta, d_flags, reinterpret_cast<unsigned long long*>(cb->d_enc), cb->d_ctx,
		                                                   uint32_t(kEncCtxMaxRows));
	}
	count_launch(2);
	MH_CUDA(cudaGetLastError());
	cb->order = order;
	cb->device_built = true;
	cb->has_box = false;
	cb->ctx_rows = 0;   // unknown on the host: the encoder reads them from d_meta
	cb->max_bits = 0;
	return MH_OK;
}

}  // namespace mh
