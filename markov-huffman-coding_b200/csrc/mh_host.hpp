// mh_host.hpp — host-side coding tables for the B200 Markov-Huffman codec.
//
// The 256 tiny Huffman trees stay on the CPU (microsecond work) but must come out bit-identical to the
// reference's: same heap tie-breaking (src/min_pq.tpp), same int32 weight arithmetic (src/tree.h:14,20),
// same "shallower subtree goes left" swap (src/huffman.cpp:147-149), same single-symbol fake root
// (src/huffman.cpp:154-162), same pre-order code assignment (src/huffman.cpp:97-123).
// Trees are index-based arenas (no pointers) so they flatten directly into the device tables.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace mh {

constexpr int kNoChild = -1;
constexpr int kMaxCodeBitsDevice = 56;

struct TreeNode {
	int left = kNoChild;
	int right = kNoChild;
	bool internal = false;
	uint8_t symbol = 0;
	int32_t weight = 0;   // int32 on purpose: wraps like the reference's `int weight`
	int32_t height = 0;
	int32_t depth = -1;
};

// A codeword of up to 255 bits, MSB-first (bit i of the codeword is bit 7 - i%8 of bytes[i/8]).
struct Codeword {
	int length = 0;
	uint64_t value = 0;   // the codeword as an integer (MSB-first reading); meaningful while length <= 64
	std::array<uint8_t, 32> bytes{};
	void append(int bit);
	void drop_last();
	int bit(int i) const { return (bytes[i >> 3] >> (7 - (i & 7))) & 1; }
};

// One Huffman tree with its derived encode table and 8-bit decode LUT.
class CodeTree {
public:
	std::vector<TreeNode> nodes;
	int root = kNoChild;
	// Derived tables exist only for non-empty trees (a Markov table has 256 trees, most of them empty for text).
	struct Derived {
		std::array<Codeword, 256> code;   // code[c].length == 0: no codeword
		std::array<int, 256> lut;         // node index per 8-bit window, kNoChild = null
	};
	std::unique_ptr<Derived> derived;

	bool empty() const { return root == kNoChild; }
	const Codeword& code(int c) const { static const Codeword none; return derived ? derived->code[c] : none; }
	int lut(int window) const { return derived ? derived->lut[window] : kNoChild; }
	int max_code_bits() const;

	// counts are the reference's ints (already truncated to int32 by the caller)
	void build_from_counts(const int32_t* counts256);
	// called after the shape exists (built or loaded): assigns depths, codewords and the LUT
	void derive_codes();
};

class BitSink;    // MSB-first bit writer
class BitSource;  // MSB-first bit reader

// The reference's i_coding_provider state: one tree (-h) or 256 trees (Markov).
class CodingTable {
public:
	int order = 1;                  // get_type()
	std::vector<CodeTree> trees;    // size 1 or 256

	static int from_counts(const uint64_t* counts, int order, CodingTable& out);       // returns mh_status
	static int from_bytes(const uint8_t* bytes, size_t n, CodingTable& out);           // returns mh_status
	std::vector<uint8_t> serialize() const;
	const CodeTree& tree_for(int prev) const { return trees[order ? (prev & 255) : 0]; }
	int max_code_bits() const;
	std::string debug_dump() const;   // -g: print_table() + print_tree()

	// ---- flat device images -------------------------------------------------------------------------
	// enc[ctx*256 + c] = (len << 56) | code (right-aligned); returns MH_ERR_CODE_TOO_LONG if any len > 56
	int flatten_codebook(uint64_t* enc /* [trees.size() * 256] */) const;
	// Smallest byte range [lo, lo + r) holding every symbol that has a codeword and every non-empty context.
	void live_range(uint32_t& lo, uint32_t& r) const;
	// box[(prev - lo) * (r + 1) + (c - lo)] (order 1) or box[c] (order 0) = len << 27 | code; needs max bits
	// <= 27. Row r and column r are a zero border that out-of-range bytes are clamped onto.
	// `box` must hold (r + 1) * (r + 1) (order 1) or 256 (order 0) entries.
	void flatten_box(uint32_t lo, uint32_t r, uint32_t* box) const;
	// Context-row table for the encoder: one 256-entry row per NON-EMPTY context (ranked in ascending byte order)
	// plus a final null row. entry = len << 27 | next_row << 16 | code, where next_row is the row of the byte as
	// the NEXT symbol's context (the null row if that context has no tree). The null row has len 0 everywhere: a
	// symbol without a codeword is dropped, like the reference's zero-length descriptor (src/coding.cpp:72). A
	// codeword longer than 16 bits is stored as the marker length 31 (the encoder reads it from the wide table). Returns the number of rows (live contexts + 1), or 0 if there are more than `max_rows`.
	uint32_t flatten_ctx(uint32_t* table /* [max_rows * 256] */, uint32_t max_rows) const;
	// lut[ctx*256 + w]: leaf : symbol << 8 | length (1..8)
	//                  deep : node index within the context << 7 | 0x10   (internal node at depth 8)
	//                  null : ' ' << 8 | 0x20 | 1                         (speculation-safe; an error if verified)
	//                  deep + ext : ext row << 7 | 0x40 | 0x10            (the first kExtRows deep nodes)
	// walk[ctx*512 + node] = left << 16 | right, child = 0x8000|symbol for a leaf, else node index
	// ext[row*256 + w2], indexed by the NEXT 8 stream bits after a deep entry: leaf within 16 bits:
	//                  symbol << 8 | extra bits (1..8); else node index (depth 16) << 7 | 0x10 -> bit-by-bit walk.
	//                  Same result as the reference's walk (src/coding.cpp:129-149), one lookup instead of <= 8 steps.
	// Returns the number of ext rows used.
	uint32_t flatten_dectable(uint16_t* lut /* [trees.size() * 256] */, uint32_t* walk /* [trees.size() * 512] */,
	                          uint16_t* ext /* [kExtRows * 256] */) const;
	// Pair table for the decoder: one 256-entry u32 row per NON-EMPTY context (ranked in ascending byte order), a
	// null row, then PREFIX rows, all indexed by the next 8 stream bits like the reference's LUT
	// (src/huffman.cpp:110-123). An entry resolves up to TWO symbols when both codewords fit in the 8 bits:
	//   [5:0]  bits consumed (1..8; bits 4 and 5 are the flags below, so four entries add up to <= 32 here)
	//   [9:6]  symbols produced (0, 1 or 2; four entries add up to <= 8 here)
	//   [15:10] next row: the last symbol as the next context (the null row if that context has no tree)
	//   [23:16] first symbol, [31:24] second symbol (0 when there is none)
	// A codeword longer than 8 bits is an internal node at depth 8 (src/coding.cpp:129-149). When every codeword
	// below that node ends within 8 more bits the node gets a prefix row: its entry in the context row consumes the
	// 8 bits, produces nothing and names the prefix row, whose entries finish the symbol — no branch in the decoder.
	//   deep (kLutDeep): a longer codeword, or no row left; null (kLutNull): no such table entry — both are decoded
	//   one symbol at a time through the u16 LUT / walk table.
	// `maps` receives rank[256] (row of each byte as a context), live[kPairMaxRows] (context byte of each context
	// row; the null row reports a context without a tree) and len1[ctx rows * 256] (length of the FIRST codeword of
	// each context-row entry, 0 for deep / null). Returns the total number of rows (<= max_rows <= kPairMaxRows), 0 if
	// the context rows alone do not fit; *ctx_rows_out = number of context rows (the null row's index).
	uint32_t flatten_pairlut(uint32_t* table /* [max_rows * 256] */, uint8_t* maps /* [256 + kPairMaxRows * 257] */, uint32_t max_rows,
	                         uint32_t* ctx_rows_out) const;
};

constexpr uint16_t kLutDeep = 0x10;
constexpr uint16_t kLutNull = 0x20;
constexpr uint16_t kLutExt = 0x40;
constexpr uint32_t kExtRows = 512;
constexpr uint32_t kWalkLeaf = 0x8000;
constexpr uint32_t kPairMaxRows = 64;   // 6-bit row field

}  // namespace mh
