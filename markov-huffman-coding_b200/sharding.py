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


# ---------------------------------------------------------------------------------------------------------
# Decode of one stream over several GPUs by bit ranges, with the seam handshake (torch.distributed + NCCL).
# ---------------------------------------------------------------------------------------------------------
HALO_HEAD = 64      # bytes a shard lends its predecessor so the predecessor's last codeword can complete
LOCAL_PAD = 4096    # bytes in front of a shard's payload inside its local buffer (room for the warm-up halo)


class ShardedDecoder:
    """Rank r holds its payload shard in `d_local[LOCAL_PAD:]` (first bit at bit `bit_base & 7`, exactly where
    mh_gpu_encode put it). decode():
      1. all-gather the halos (each rank's last `tail_bytes` and first HALO_HEAD bytes) and splice them around the
         local payload, OR-merging the byte two neighbours share;
      2. mh_gpu_decode_shard: rank 0 starts exactly; every other rank starts `warm_bits` bits before its
         first bit from a guessed state and converges over that warm-up;
      3. all-gather (symbol count, seam words). Rank r is consistent when the state it reached at its first bit
         equals the state rank r-1 ended in; a rank that is not decodes again from that exact state. Repeat until
         every seam agrees (normally zero repeats);
      4. exclusive scan of the counts -> every rank's offset in the decoded stream.
    Only tiny messages cross NVLink: (tail_bytes + HALO_HEAD) bytes and 24 bytes per rank."""

    def __init__(self, mh, dist, torch, rank, world, order, device):
        self.mh, self.dist, self.torch = mh, dist, torch
        self.rank, self.world, self.order = rank, world, order
        self.warm_bits = mh.DECODE_WARM_UNIT                 # several synchronisation distances
        self.tail_bytes = self.warm_bits // 8 + 8
        assert self.tail_bytes + 8 <= LOCAL_PAD
        self.my_halo = torch.zeros(self.tail_bytes + HALO_HEAD, dtype=torch.uint8, device=device)
        self.halos = torch.zeros(world * (self.tail_bytes + HALO_HEAD), dtype=torch.uint8, device=device)
        self.my_seam = torch.zeros(6, dtype=torch.int64, device=device)   # d_result[0..3], exact-start marker, spare
        self.seams = torch.zeros(6 * world, dtype=torch.int64, device=device)
        self.h_seams = torch.zeros(6 * world, dtype=torch.int64).pin_memory()
        self.rounds = 0

    def decode(self, d_local, bit_bases, bits, dectab, d_out, out_capacity, d_result, ws, stream):
        torch, dist, mh, r = self.torch, self.dist, self.mh, self.rank
        T, H = self.tail_bytes, HALO_HEAD
        phase = int(bit_bases[r]) & 7
        nbytes = (phase + int(bits[r]) + 7) // 8
        pay = d_local[LOCAL_PAD:]
        # ---- 1. halo exchange ----
        self.my_halo[:T] = pay[nbytes - T:nbytes]
        self.my_halo[T:] = pay[:H]
        dist.all_gather_into_tensor(self.halos, self.my_halo)
        halos = self.halos.view(self.world, T + H)
        if r > 0:
            tail = halos[r - 1, :T]
            if phase:      # the predecessor's last byte is my first byte: OR the seam
                d_local[LOCAL_PAD - T + 1:LOCAL_PAD] = tail[:-1]
                pay[0] |= tail[-1]
            else:
                d_local[LOCAL_PAD - T:LOCAL_PAD] = tail
        if r + 1 < self.world:
            head = halos[r + 1, T:]
            if int(bit_bases[r + 1]) & 7:
                pay[nbytes - 1] |= head[0]
                pay[nbytes:nbytes + H - 1] = head[1:]
            else:
                pay[nbytes:nbytes + H] = head
        else:
            pay[nbytes:nbytes + H] = 0
        buf_end = LOCAL_PAD + nbytes + H
        # ---- 2. speculative decode of my bit range ----
        own_bit = LOCAL_PAD * 8 + phase                      # my first bit, in local-buffer bit coordinates
        exact, prev0, start = (r == 0), 0x20, own_bit
        warm = 0 if r == 0 else self.warm_bits
        self.rounds = 0
        started_from = -1                                    # the predecessor end state I started from exactly, if any
        while True:
            origin = start - warm
            off = (origin // 32) * 4
            n_bits = own_bit + int(bits[r]) - origin
            mh.gpu_decode_shard(d_local.data_ptr() + off, origin % 32, n_bits, buf_end - off, exact, prev0, warm, r == self.world - 1,
                                dectab, d_out.data_ptr(), out_capacity, d_result.data_ptr(), ws, stream)
            # ---- 3. handshake ----
            self.my_seam[:4] = d_result
            self.my_seam[4] = started_from
            dist.all_gather_into_tensor(self.seams, self.my_seam)
            self.h_seams.copy_(self.seams, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            s = self.h_seams.numpy().reshape(self.world, 6)
            assert not s[:, 1].any(), "a shard reported status %s" % s[:, 1].tolist()
            words = s[:, 3].view(np.uint64) if s[:, 3].flags.c_contiguous else np.ascontiguousarray(s[:, 3]).view(np.uint64)
            views, ends = (words >> np.uint64(32)).astype(np.int64), (words & np.uint64(0xFFFFFFFF)).astype(np.int64)
            ok = [g == 0 or (s[g, 4] == ends[g - 1] if s[g, 4] >= 0 else views[g] == ends[g - 1]) for g in range(self.world)]
            if all(ok):
                break
            self.rounds += 1
            assert self.rounds <= self.world + 1, "seam handshake did not converge"
            if not ok[r]:      # decode again, this time from the state my predecessor really ended in
                e = int(ends[r - 1])
                exact, prev0, warm, start, started_from = True, e & 255, 0, own_bit + (e >> 8), e
        # ---- 4. output offsets ----
        counts = s[:, 0]
        self.out_offset = int(counts[:r].sum())
        self.out_count = int(counts[r])
        return self.out_count
