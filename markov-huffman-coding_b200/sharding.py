"""Host-side arithmetic for sharding ONE stream over several GPUs by byte range: the reference
arithmetic that tests/test_sharding.py runs under gloo on CPU and that tests/test_sharded_lib.py merges the library's
shards with. The product path (histogram gather, bit offsets, halo exchange, handshake) is csrc/mh_shard.cu.

Nothing here touches the data: every quantity comes from the per-shard histograms that the histogram kernel
already produced.
  * shard g's histogram and encoder are seeded with the last byte of shard g-1 (prev0), so seam pairs are counted
    and coded exactly once;
  * the global histogram is the sum of the all-gathered local ones, so every rank builds identical tables;
  * shard g's payload size is sum(local_count[g] * code_length), hence every rank knows every shard's global bit
    offset (bit_base) without another exchange;
  * a shard's payload buffer starts at bit (bit_base & 7) of its first byte; assembling the stream ORs the byte
    that two neighbouring shards share.
"""
import numpy as np


def global_counts(all_counts):
    """all_counts: uint64 [world, bins] (all-gathered local histograms) -> uint64 [bins]."""
    a = np.asarray(all_counts, dtype=np.uint64)
    if a.ndim == 1 or a.shape[0] == 1:
        return np.ascontiguousarray(a.reshape(-1))
    out = a[0] + a[1]                       # row-wise adds run along the contiguous axis (a [world, bins] reduction over
    for r in range(2, a.shape[0]):          # axis 0 is ~5x slower in numpy, and this sits between histogram and encoder)
        out += a[r]
    return out


def shard_bits(all_counts, code_lengths, total=None):
    """Payload bits of every shard: uint64 [world]. Only the (prev, c) pairs that have a codeword contribute.
    (`total` is accepted for symmetry with global_counts; the code lengths already say which pairs are live.)"""
    a = np.asarray(all_counts, dtype=np.uint64)
    lens = np.asarray(code_lengths)
    if lens.size == 65536:                  # order 1: look for the live pairs only in the rows of live contexts (a few
        l2 = lens.reshape(256, 256)         # dozen for text) instead of scanning 65536 entries
        rows = np.flatnonzero(l2.any(axis=1))
        sub = l2[rows]
        r, c = np.nonzero(sub)
        return np.take(a, rows[r] * 256 + c, axis=1) @ sub[r, c].astype(np.uint64)
    live = np.flatnonzero(lens)
    return np.take(a, live, axis=1) @ lens[live].astype(np.uint64)


def fix_seam_pairs(all_counts, first_bytes, last_bytes, guess=0x20):
    """Every shard counted its first byte as following `guess`; move that one count to the pair it really forms with
    the previous shard's last byte (order-1 histograms only). all_counts: uint64 [world, 65536], modified in place."""
    for g in range(1, len(first_bytes)):
        f, true_prev = int(first_bytes[g]), int(last_bytes[g - 1])
        all_counts[g, 256 * guess + f] -= 1
        all_counts[g, 256 * true_prev + f] += 1
    return all_counts


def fix_seam_total(total, first_bytes, last_bytes, guess=0x20):
    """The same correction as fix_seam_pairs, applied to the SUM of the local histograms (uint64 [65536], in place):
    used when the sum was reduced on the GPU before the seam bytes were known."""
    for g in range(1, len(first_bytes)):
        f, true_prev = int(first_bytes[g]), int(last_bytes[g - 1])
        total[256 * guess + f] -= 1
        total[256 * true_prev + f] += 1
    return total


def shard_bit_bases(all_counts, code_lengths, total=None):
    """(bit_base [world], bits [world]): exclusive scan of the shard payload sizes."""
    bits = shard_bits(all_counts, code_lengths, total)
    base = np.zeros_like(bits)
    if len(bits) > 1:
        base[1:] = np.cumsum(bits[:-1], dtype=np.uint64)
    return base, bits


def merge_payload_shards(shards):
    """shards: iterable of (payload bytes as produced by mh_gpu_encode, bit_base, n_bits), in rank order.
    Returns the single payload (no header byte) the unsharded encoder would have produced."""
    shards = list(shards)
    total_bits = int(shards[-1][1]) + int(shards[-1][2]) if shards else 0
    out = np.zeros((total_bits + 7) // 8, dtype=np.uint8)
    for payload, bit_base, n_bits in shards:
        bit_base, n_bits = int(bit_base), int(n_bits)
        if n_bits == 0:
            continue
        first = bit_base // 8
        nbytes = ((bit_base & 7) + n_bits + 7) // 8
        buf = np.frombuffer(payload, dtype=np.uint8)[:nbytes].copy()
        # keep only this shard's own bits in its first and last byte, then OR into place
        buf[0] &= 0xFF >> (bit_base & 7)
        tail = (bit_base + n_bits) & 7
        if tail:
            buf[-1] &= (0xFF << (8 - tail)) & 0xFF
        out[first:first + nbytes] |= buf
    return out.tobytes()


def stream_header(order, total_bits):
    """Header byte 0 0 1 1 E R R R (reference src/coding.cpp:88)."""
    return 0x30 | ((~order & 1) << 3) | ((8 - total_bits % 8) % 8)


# The decode side of the multi-GPU path (halo exchange, speculative start with warm-up, seam handshake) lives in the
# library since round 2: csrc/mh_shard.cu, mh_sharded_decompress (bound as Comm.decompress in __init__.py). It refuses
# shards shorter than the warm-up in speculative mode and handles empty shards through the exact layout.
