"""Host-side arithmetic for sharding ONE stream over several GPUs by byte range.

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
    return np.ascontiguousarray(np.asarray(all_counts, dtype=np.uint64).sum(axis=0, dtype=np.uint64))


def shard_bits(all_counts, code_lengths):
    """Payload bits of every shard: uint64 [world]."""
    return (np.asarray(all_counts, dtype=np.uint64) * np.asarray(code_lengths, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)


def shard_bit_bases(all_counts, code_lengths):
    """(bit_base [world], bits [world]): exclusive scan of the shard payload sizes."""
    bits = shard_bits(all_counts, code_lengths)
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
