// mh_host.cpp — host-side coding tables (see mh_host.hpp). Reference lines are cited where behaviour is a
// contract; the code is an index-arena restatement, not a translation.
#include "mh_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/mh_gpu.h"

namespace mh {

// ------------------------------------------------------------------------------------------------------
// Codeword: MSB-first bit string (reference: encoding_descriptor, src/coding.h:9-16, src/coding.cpp:9-27)
// ------------------------------------------------------------------------------------------------------
void Codeword::append(int bit) {
	if(bit) bytes[length >> 3] |= uint8_t(0x80u >> (length & 7));
	value = (value << 1) | uint64_t(bit & 1);
	++length;
}

void Codeword::drop_last() {
	--length;
	bytes[length >> 3] &= uint8_t(~(0x80u >> (length & 7)));
	value >>= 1;
}

// ------------------------------------------------------------------------------------------------------
// Binary min-heap over (weight, node). The comparison pattern below IS the format: it decides which of two
// equal-weight subtrees is merged first and therefore every tie in the tree shape (src/min_pq.tpp:29-52).
//   sift-up   : move while parent.weight >  self.weight            (strict)
//   sift-down : candidate = right child only if right.weight < left.weight (strict), else left;
//               swap only if candidate.weight < self.weight        (strict)
// ------------------------------------------------------------------------------------------------------
namespace {

class WeightHeap {
	struct Slot { int32_t weight; int node; };
	std::vector<Slot> a_;
public:
	size_t size() const { return a_.size(); }
	void reserve(size_t n) { a_.reserve(n); }
	void push(int32_t weight, int node) {
		a_.push_back({weight, node});
		size_t i = a_.size() - 1;
		while(i > 0) {
			size_t up = (i - 1) >> 1;
			if(!(a_[up].weight > a_[i].weight)) break;
			std::swap(a_[up], a_[i]);
			i = up;
		}
	}
	int pop() {
		int top = a_.front().node;
		a_.front() = a_.back();
		a_.pop_back();
		const size_t n = a_.size();
		size_t i = 0;
		for(;;) {
			size_t l = 2 * i + 1, r = l + 1;
			size_t pick = (r < n && a_[r].weight < a_[l].weight) ? r : l;
			if(pick >= n || !(a_[pick].weight < a_[i].weight)) break;
			std::swap(a_[pick], a_[i]);
			i = pick;
		}
		return top;
	}
};

inline int32_t wrap_add(int32_t x, int32_t y) { return int32_t(uint32_t(x) + uint32_t(y)); }

}  // namespace

// ------------------------------------------------------------------------------------------------------
// CodeTree
// ------------------------------------------------------------------------------------------------------
void CodeTree::build_from_counts(const int32_t* counts) {
	nodes.clear();
	root = kNoChild;
	derived.reset();
	int live = 0;
	for(int s = 0; s < 256; ++s) live += counts[s] != 0;
	if(live == 0) return;   // empty table (most contexts of a text): nothing to allocate
	nodes.reserve(size_t(2 * live + 2));
	WeightHeap heap;
	heap.reserve(size_t(live));
	for(int s = 0; s < 256; ++s) {          // ascending symbol order fixes the initial heap layout (src/huffman.cpp:134-138)
		if(counts[s] == 0) continue;        // the reference tests `if(counts[i])`: a wrapped-negative count is still a leaf
		TreeNode leaf;
		leaf.symbol = uint8_t(s);
		leaf.weight = counts[s];
		nodes.push_back(leaf);
		heap.push(counts[s], int(nodes.size()) - 1);
	}
	if(heap.size() == 0) return;   // empty table: a freshly constructed CodeTree is already clean
	while(heap.size() > 1) {
		int a = heap.pop();
		int b = heap.pop();
		if(nodes[a].height > nodes[b].height) std::swap(a, b);   // shallower subtree on the left (src/huffman.cpp:147-149)
		TreeNode parent;
		parent.internal = true;
		parent.left = a;
		parent.right = b;
		parent.weight = wrap_add(nodes[a].weight, nodes[b].weight);
		parent.height = std::max(nodes[a].height, nodes[b].height) + 1;
		nodes.push_back(parent);
		heap.push(parent.weight, int(nodes.size()) - 1);
	}
	root = heap.pop();
	if(!nodes[root].internal) {
		// One live symbol: the reference turns the lone leaf into an internal root over two copies of itself, so
		// the symbol gets the 1-bit code "1" (the right copy is visited last) and the tree serialises as
		// 0, 1 vvvvvvvv, 1 vvvvvvvv (src/huffman.cpp:154-162).
		TreeNode copy = nodes[root];
		nodes.push_back(copy);
		nodes.push_back(copy);
		nodes[root].internal = true;
		nodes[root].left = int(nodes.size()) - 2;
		nodes[root].right = int(nodes.size()) - 1;
		nodes[root].height = 1;
	}
	derive_codes();
}

void CodeTree::derive_codes() {
	derived.reset();
	if(root == kNoChild) return;
	derived.reset(new Derived());
	auto& code = derived->code;
	auto& lut = derived->lut;
	lut.fill(kNoChild);
	// Iterative pre-order walk, left edge = 0, right edge = 1 (src/huffman.cpp:97-123). `stage` per frame:
	// 0 = entering, 1 = left subtree done, 2 = right subtree done.
	struct Frame { int node; int stage; };
	std::vector<Frame> stack;
	Codeword path;
	stack.push_back({root, 0});
	while(!stack.empty()) {
		Frame& f = stack.back();
		TreeNode& nd = nodes[f.node];
		const int depth = int(stack.size()) - 1;
		if(!nd.internal) {
			nd.depth = depth;
			code[nd.symbol] = path;                    // a symbol reached twice keeps the LAST path (single-symbol case)
			if(depth <= 8) {                           // src/huffman.cpp:116-121
				const int first = path.bytes[0];
				for(int i = 0; i < (1 << (8 - depth)); ++i) lut[first + i] = f.node;
			}
			stack.pop_back();
			if(!stack.empty()) path.drop_last();
			continue;
		}
		if(f.stage == 0) {
			nd.depth = depth;
			f.stage = 1;
			path.append(0);
			stack.push_back({nd.left, 0});
		} else if(f.stage == 1) {
			f.stage = 2;
			path.append(1);
			stack.push_back({nd.right, 0});
		} else {
			if(depth == 8) lut[path.bytes[0]] = f.node;   // internal node exactly 8 deep owns its window (src/huffman.cpp:111-113)
			stack.pop_back();
			if(!stack.empty()) path.drop_last();
		}
	}
}

int CodeTree::max_code_bits() const {
	int m = 0;
	if(empty()) return 0;
	for(const auto& c : derived->code) m = std::max(m, c.length);
	return m;
}

// ------------------------------------------------------------------------------------------------------
// MSB-first bit I/O for the table file (the subset of src/bitbuffer.cpp the file format needs): stream bit p
// is bit 7 - p%8 of byte p/8; the writer zero-pads the last byte (src/bitbuffer.cpp:170-186).
// ------------------------------------------------------------------------------------------------------
class BitSink {
	std::vector<uint8_t>& out_;
	uint64_t nbits_ = 0;
public:
	explicit BitSink(std::vector<uint8_t>& out) : out_(out) {}
	void bit(int b) {
		if((nbits_ & 7) == 0) out_.push_back(0);
		if(b) out_.back() |= uint8_t(0x80u >> (nbits_ & 7));
		++nbits_;
	}
	void byte(uint8_t v) { for(int i = 7; i >= 0; --i) bit((v >> i) & 1); }
};

class BitSource {
	const uint8_t* p_;
	uint64_t nbits_, pos_ = 0;
	bool ok_ = true;
public:
	BitSource(const uint8_t* p, size_t nbytes) : p_(p), nbits_(uint64_t(nbytes) * 8) {}
	bool ok() const { return ok_; }
	int bit() {
		if(pos_ >= nbits_) { ok_ = false; return 0; }
		int b = (p_[pos_ >> 3] >> (7 - (pos_ & 7))) & 1;
		++pos_;
		return b;
	}
	uint8_t byte() { uint8_t v = 0; for(int i = 0; i < 8; ++i) v = uint8_t((v << 1) | bit()); return v; }
};

namespace {

// writer: internal -> 0; leaf -> 1 + symbol; left subtree, then right subtree (src/huffman.cpp:174-188)
void emit_tree(const CodeTree& t, BitSink& sink) {
	if(t.empty()) return;
	std::vector<int> todo{t.root};
	while(!todo.empty()) {
		const TreeNode& nd = t.nodes[todo.back()];
		todo.pop_back();
		if(nd.internal) {
			sink.bit(0);
			todo.push_back(nd.right);   // popped after the whole left subtree
			todo.push_back(nd.left);
		} else {
			sink.bit(1);
			sink.byte(nd.symbol);
		}
	}
}

// loader: mirror of the writer, LEFT subtree first (the reference's :170 leaves the order to the compiler; the
// writer and the format comment at src/huffman.cpp:75-81 define it). Returns false on truncation or on a shape
// the reference cannot represent (leaf-only root, depth > 255, more than 511 nodes).
bool parse_tree(BitSource& src, CodeTree& t) {
	t.nodes.clear();
	t.root = kNoChild;
	struct Pending { int node; int filled; };
	std::vector<Pending> open;    // internal nodes still waiting for children
	for(;;) {
		const int is_leaf = src.bit();
		if(!src.ok()) return false;
		if(t.nodes.size() >= 511) return false;
		TreeNode nd;
		if(is_leaf) {
			nd.symbol = src.byte();
			if(!src.ok()) return false;
		} else {
			nd.internal = true;
		}
		const int me = int(t.nodes.size());
		t.nodes.push_back(nd);
		if(open.empty()) {
			if(is_leaf) return false;           // a bare leaf as root never comes out of the writer
			t.root = me;
		} else {
			Pending& p = open.back();
			if(p.filled == 0) t.nodes[p.node].left = me; else t.nodes[p.node].right = me;
			++p.filled;
		}
		if(!is_leaf) {
			if(open.size() >= 255) return false;
			open.push_back({me, 0});
		}
		// close every internal node that now has both children, fixing heights bottom-up
		while(!open.empty() && open.back().filled == 2) {
			TreeNode& done = t.nodes[open.back().node];
			done.height = std::max(t.nodes[done.left].height, t.nodes[done.right].height) + 1;
			open.pop_back();
		}
		if(open.empty()) return true;
	}
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// CodingTable
// ------------------------------------------------------------------------------------------------------
int CodingTable::from_counts(const uint64_t* counts, int order, CodingTable& out) {
	if(order != 0 && order != 1) return MH_ERR_INVALID_ARG;
	out.order = order;
	const int ntab = order ? 256 : 1;
	out.trees.clear();
	out.trees.resize(ntab);
	for(int t = 0; t < ntab; ++t) {
		int32_t c32[256];
		const uint64_t* row = counts + size_t(t) * 256;
		uint64_t any = 0;
		uint32_t wrapped = 0;
		for(int s = 0; s < 256; ++s) {            // branch-free: most rows of a text table are all zero
			const uint64_t c = row[s];
			c32[s] = int32_t(uint32_t(c));      // the reference counts in `int` (src/main.cpp:166,174): keep the low 32 bits
			any |= c;
			wrapped |= uint32_t(c != 0) & uint32_t(uint32_t(c) == 0);
		}
		if(wrapped) return MH_ERR_COUNT_WRAPPED;   // the reference would silently lose a live symbol
		if(any) out.trees[t].build_from_counts(c32);    // src/markov_huffman.cpp:9-13 (an all-zero row leaves the tree empty)
	}
	return MH_OK;
}

int CodingTable::from_bytes(const uint8_t* bytes, size_t n, CodingTable& out) {
	if(n == 0 || bytes == nullptr) return MH_ERR_BAD_TABLE;   // 0-byte file = empty -h table; the reference cannot load it either
	BitSource src(bytes, n);
	const bool markov = (bytes[0] & 0x80) != 0;               // peek_bit (src/main.cpp:147)
	out.order = markov ? 1 : 0;
	out.trees.clear();
	out.trees.resize(markov ? 256 : 1);
	if(markov) {
		src.bit();                                            // kind marker (src/markov_huffman.cpp:17)
		for(int p = 0; p < 256; ++p) {
			const int present = src.bit();                    // src/markov_huffman.cpp:20
			if(!src.ok()) return MH_ERR_BAD_TABLE;
			if(present && !parse_tree(src, out.trees[p])) return MH_ERR_BAD_TABLE;
		}
	} else {
		if(!parse_tree(src, out.trees[0])) return MH_ERR_BAD_TABLE;
	}
	for(auto& t : out.trees) t.derive_codes();                // src/huffman.cpp:22-25
	return MH_OK;
}

std::vector<uint8_t> CodingTable::serialize() const {
	std::vector<uint8_t> out;
	BitSink sink(out);
	if(order) {
		sink.bit(1);                                          // src/markov_huffman.cpp:81
		for(const auto& t : trees) {
			sink.bit(t.empty() ? 0 : 1);                      // :83
			emit_tree(t, sink);
		}
	} else {
		emit_tree(trees[0], sink);                            // src/huffman.cpp:83-85
	}
	return out;
}

int CodingTable::max_code_bits() const {
	int m = 0;
	for(const auto& t : trees) m = std::max(m, t.max_code_bits());
	return m;
}

int CodingTable::flatten_codebook(uint64_t* enc) const {
	std::fill(enc, enc + trees.size() * 256, uint64_t(0));
	for(size_t t = 0; t < trees.size(); ++t) {
		if(trees[t].empty()) continue;
		for(int c = 0; c < 256; ++c) {
			const Codeword& cw = trees[t].code(c);
			if(cw.length == 0) continue;
			if(cw.length > kMaxCodeBitsDevice) return MH_ERR_CODE_TOO_LONG;
			enc[t * 256 + c] = (uint64_t(cw.length) << 56) | cw.value;
		}
	}
	return MH_OK;
}

void CodingTable::live_range(uint32_t& lo, uint32_t& r) const {
	int first = 256, last = -1;
	auto touch = [&](int s) { first = std::min(first, s); last = std::max(last, s); };
	for(size_t t = 0; t < trees.size(); ++t) {
		if(trees[t].empty()) continue;
		if(order) touch(int(t));
		for(int c = 0; c < 256; ++c)
			if(trees[t].code(c).length) touch(c);
	}
	if(last < 0) { lo = 0; r = 1; return; }
	lo = uint32_t(first);
	r = uint32_t(last - first + 1);
}

void CodingTable::flatten_box(uint32_t lo, uint32_t r, uint32_t* box) const {
	auto entry = [](const Codeword& cw) -> uint32_t {
		return cw.length ? (uint32_t(cw.length) << 27) | uint32_t(cw.value) : 0u;
	};
	if(!order) {
		for(int c = 0; c < 256; ++c) box[c] = entry(trees[0].code(c));
		return;
	}
	const uint32_t pitch = r + 1;
	std::fill(box, box + size_t(pitch) * pitch, 0u);
	for(uint32_t p = 0; p < r; ++p) {
		if(trees[lo + p].empty()) continue;
		for(uint32_t c = 0; c < r; ++c) box[p * pitch + c] = entry(trees[lo + p].code(lo + c));
	}
}

uint32_t CodingTable::flatten_ctx(uint32_t* table, uint32_t max_rows) const {
	int rank[256];
	uint32_t live = 0;
	for(int p = 0; p < 256; ++p) {
		const bool has = order ? !trees[p].empty() : (p == 0 && !trees[0].empty());
		rank[p] = has ? int(live++) : -1;
	}
	const uint32_t rows = live + 1;   // + the null row
	if(rows > max_rows) return 0;
	auto next_of = [&](int c) -> uint32_t { return order ? (rank[c] >= 0 ? uint32_t(rank[c]) : live) : 0u; };
	for(int p = 0; p < 256; ++p) {
		if(rank[p] < 0) continue;
		const CodeTree& tr = trees[order ? p : 0];
		uint32_t* row = table + size_t(rank[p]) * 256;
		for(int c = 0; c < 256; ++c) {
			const Codeword& cw = tr.code(c);
			// codewords longer than 16 bits (rare by construction) carry the marker length 31: the encoder then takes
			// the symbol from the wide table instead
			row[c] = cw.length > 16 ? (31u << 27) | (next_of(c) << 16)
			                        : (uint32_t(cw.length) << 27) | (next_of(c) << 16) | (cw.length ? uint32_t(cw.value) : 0u);
		}
	}
	uint32_t* null_row = table + size_t(live) * 256;
	for(int c = 0; c < 256; ++c) null_row[c] = next_of(c) << 16;
	return rows;
}

uint32_t CodingTable::flatten_dectable(uint16_t* lut, uint32_t* walk, uint16_t* ext) const {
	std::fill(lut, lut + trees.size() * 256, uint16_t((' ' << 8) | kLutNull | 1u));
	std::fill(walk, walk + trees.size() * 512, uint32_t(0));
	uint32_t rows = 0;
	for(size_t t = 0; t < trees.size(); ++t) {
		const CodeTree& tr = trees[t];
		if(tr.empty()) continue;
		for(int w = 0; w < 256; ++w) {
			const int n = tr.lut(w);
			if(n == kNoChild) continue;
			const TreeNode& nd = tr.nodes[n];
			if(!nd.internal) {
				lut[t * 256 + w] = uint16_t((nd.symbol << 8) | nd.depth);
			} else if(rows < kExtRows) {
				// second-level row for this depth-8 node: follow the next 8 bits, MSB first
				uint16_t* row = ext + size_t(rows) * 256;
				for(int w2 = 0; w2 < 256; ++w2) {
					int cur = n, d = 0;
					while(tr.nodes[cur].internal && d < 8) {
						cur = ((w2 >> (7 - d)) & 1) ? tr.nodes[cur].right : tr.nodes[cur].left;
						++d;
					}
					row[w2] = tr.nodes[cur].internal ? uint16_t((cur << 7) | kLutDeep) : uint16_t((tr.nodes[cur].symbol << 8) | d);
				}
				lut[t * 256 + w] = uint16_t((rows << 7) | kLutExt | kLutDeep);
				++rows;
			} else {
				lut[t * 256 + w] = uint16_t((n << 7) | kLutDeep);
			}
		}
		for(size_t n = 0; n < tr.nodes.size(); ++n) {
			const TreeNode& nd = tr.nodes[n];
			if(!nd.internal) continue;
			auto child = [&](int c) -> uint32_t {
				const TreeNode& ch = tr.nodes[c];
				return ch.internal ? uint32_t(c) : (kWalkLeaf | ch.symbol);
			};
			walk[t * 512 + n] = (child(nd.left) << 16) | child(nd.right);
		}
	}
	return rows;
}

uint32_t CodingTable::flatten_pairlut(uint32_t* table, uint8_t* maps, uint32_t max_rows, uint32_t* ctx_rows_out) const {
	int rank[256];
	uint32_t live = 0;
	int dead = -1;
	for(int p = 0; p < 256; ++p) {
		const bool has = order ? !trees[p].empty() : (p == 0 && !trees[0].empty());
		rank[p] = has ? int(live++) : -1;
		if(!has && dead < 0 && (order || p > 0)) dead = p;
	}
	if(max_rows > kPairMaxRows) max_rows = kPairMaxRows;
	uint32_t rows = live + 1;   // context rows + the null row; prefix rows follow
	if(rows > max_rows || live == 0) return 0;
	auto next_of = [&](int c) -> uint32_t { return order ? (rank[c] >= 0 ? uint32_t(rank[c]) : live) : 0u; };
	uint8_t* len1 = maps + 256 + kPairMaxRows;
	for(int p = 0; p < 256; ++p) maps[p] = uint8_t(order ? (rank[p] >= 0 ? uint32_t(rank[p]) : live) : 0u);
	for(uint32_t r = 0; r < kPairMaxRows; ++r) maps[256 + r] = uint8_t(dead < 0 ? 0 : dead);
	struct DeepNode { uint32_t weight; int ctx, window, node; };
	std::vector<DeepNode> deep;
	for(int p = 0; p < 256; ++p) {
		if(rank[p] < 0) continue;
		maps[256 + rank[p]] = uint8_t(p);
		const CodeTree& tr = trees[order ? p : 0];
		uint32_t* row = table + size_t(rank[p]) * 256;
		uint8_t* l1 = len1 + size_t(rank[p]) * 256;
		for(int w = 0; w < 256; ++w) {
			l1[w] = 0;
			const int n1 = tr.lut(w);
			if(n1 == kNoChild) { row[w] = kLutNull; continue; }
			const TreeNode& a = tr.nodes[n1];
			if(a.internal) {   // depth-8 internal node: a prefix row if one is left (below), else the slow path
				row[w] = kLutDeep;
				deep.push_back({uint32_t(a.weight), p, w, n1});
				continue;
			}
			const uint32_t d1 = uint32_t(a.depth);
			l1[w] = uint8_t(d1);
			uint32_t e = d1 | (1u << 6) | (next_of(a.symbol) << 10) | (uint32_t(a.symbol) << 16);
			// second symbol: the 8 - d1 bits left in the window must decide it completely
			const CodeTree& tr2 = trees[order ? a.symbol : 0];
			if(d1 < 8 && !tr2.empty()) {
				const int n2 = tr2.lut((w << d1) & 255);
				if(n2 != kNoChild) {
					const TreeNode& b = tr2.nodes[n2];
					if(!b.internal && uint32_t(b.depth) <= 8 - d1)
						e = (d1 + uint32_t(b.depth)) | (2u << 6) | (next_of(b.symbol) << 10) | (uint32_t(a.symbol) << 16) | (uint32_t(b.symbol) << 24);
				}
			}
			row[w] = e;
		}
	}
	// Prefix rows go to the heaviest depth-8 nodes first (the node weight is the number of symbols coded below it;
	// a table loaded from a file has no weights and takes them in table order). A codeword that ends within the row's 8
	// bits is an ordinary one-symbol entry; one that does not (17 bits and more: rare even in a Fibonacci-shaped tree,
	// 2^-16 of the symbols) is a deep flag that names the tree node reached ([31:24], ninth bit in [6]) and the context
	// ([23:16]); its next-row field stays 0, so whatever the decoder looks up speculatively behind it stays inside the
	// table. The decoder takes the row's 8 bits and walks on from that node. (Round 1 gave rows only to nodes whose codewords all
	// end within 16 bits: a Fibonacci-shaped tree then got none, and 0.4 % of its symbols took the slow path.)
	std::stable_sort(deep.begin(), deep.end(), [](const DeepNode& x, const DeepNode& y) { return x.weight > y.weight; });
	for(const DeepNode& dn : deep) {
		if(rows >= max_rows) break;
		const CodeTree& tr = trees[order ? dn.ctx : 0];
		uint32_t* ext = table + size_t(rows) * 256;
		for(int w2 = 0; w2 < 256; ++w2) {
			int cur = dn.node, d = 0;
			while(tr.nodes[cur].internal && d < 8) {
				cur = ((w2 >> (7 - d)) & 1) ? tr.nodes[cur].right : tr.nodes[cur].left;
				++d;
			}
			if(tr.nodes[cur].internal) ext[w2] = kLutDeep | ((uint32_t(cur) >> 8) << 6) | (uint32_t(dn.ctx) << 16) | ((uint32_t(cur) & 255u) << 24);
			else ext[w2] = uint32_t(d) | (1u << 6) | (next_of(tr.nodes[cur].symbol) << 10) | (uint32_t(tr.nodes[cur].symbol) << 16);
		}
		table[size_t(rank[dn.ctx]) * 256 + dn.window] = 8u | (rows << 10);   // 8 bits, no symbol yet, continue in the prefix row
		++rows;
	}
	uint32_t* null_row = table + size_t(live) * 256;
	for(int w = 0; w < 256; ++w) null_row[w] = kLutNull;
	if(ctx_rows_out) *ctx_rows_out = live;
	return rows;
}

// ------------------------------------------------------------------------------------------------------
// -g debug dump: byte-identical to print_table() + print_tree() on stdout
// (src/huffman.cpp:52-69, src/markov_huffman.cpp:31-50, src/tree.cpp:11-53, src/utils.cpp:18-42).
// ------------------------------------------------------------------------------------------------------
namespace {

std::string printable(uint8_t c) {     // charv(): the escapes are doubled because the text is meant for Graphviz
	switch(c) {
		case ' ': return "\\\\sp";
		case '\t': return "\\\\t";
		case '\r': return "\\\\r";
		case '\n': return "\\\\n";
		case '"': return "\\\"";
		case '\'': return "\\'";
		case '\\': return "\\\\";
	}
	if(c > 32 && c < 127) return std::string(1, char(c));
	char buf[16];
	snprintf(buf, sizeof buf, "\\\\%x", unsigned(c));
	return buf;
}

void dump_codes(const CodeTree& t, std::string& out) {
	out += "Table:\n";
	for(int c = 0; c < 256; ++c) {
		const Codeword& cw = t.code(c);
		if(!cw.length) continue;
		out += printable(uint8_t(c)) + " " + std::to_string(cw.length) + " ";
		for(int i = 0; i < cw.length; ++i) out += char('0' + cw.bit(i));
		out += "\n";
	}
}

// Numbers nodes in pre-order; an edge line follows the whole subtree it leads to (src/tree.cpp:32-53).
int dump_nodes(const CodeTree& t, int node, int n, std::string& out) {
	if(node == kNoChild) return -1;
	const TreeNode& nd = t.nodes[node];
	const std::string me = "\tn" + std::to_string(n);
	out += me + ";\n";
	out += me + " [label=\"" + (nd.internal ? std::string() : printable(nd.symbol)) + "\"];\n";
	int next = n + 1, last = n;
	int l = dump_nodes(t, nd.left, next, out);
	if(l != -1) { out += me + " -- n" + std::to_string(next) + ";\n"; last = l; next = l + 1; }
	int r = dump_nodes(t, nd.right, next, out);
	if(r != -1) { out += me + " -- n" + std::to_string(next) + ";\n"; last = r; }
	return last;
}

int dump_graph(const CodeTree& t, bool subgraph, int n, const std::string& label, std::string& out) {
	if(subgraph) {
		out += "subgraph clusterG" + std::to_string(n) + " {\n\tlabel=\"" + label + "\";\n\tcolor=invis;\n";
	} else {
		out += "graph G {\n";
	}
	out += "\tnodesep=0.3;\n\tranksep=0.2;\n\tnode [shape=circle, fixedsize=true];\n\tedge [arrowsize=0.8];\n";
	n = dump_nodes(t, t.root, n, out) + 1;
	out += "}\n";
	return n;
}

}  // namespace

std::string CodingTable::debug_dump() const {
	std::string out;
	if(order == 0) {
		dump_codes(trees[0], out);
		if(!trees[0].empty()) dump_graph(trees[0], false, 0, "", out);   // the reference dereferences null here (App. D9)
		return out;
	}
	for(int p = 0; p < 256; ++p)
		if(!trees[p].empty()) {
			out += "Prev '" + printable(uint8_t(p)) + "' table:\n";
			dump_codes(trees[p], out);
		}
	out += "graph G {\n\tpackmode=\"cluster\";\n";
	int n = 0;
	for(int p = 0; p < 256; ++p)
		if(!trees[p].empty()) {
			out += "/* Prev '" + printable(uint8_t(p)) + "' tree: */\n";
			n = dump_graph(trees[p], true, n, "Prev: " + printable(uint8_t(p)), out);
		}
	out += "}\n";
	return out;
}

}  // namespace mh
