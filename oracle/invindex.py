"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU restatement of build-p Part 3 and of the serialized
forms it produces, used to check lphb_build_inverted_index.  Pinned by tests/test_oracle_golden.py against the
`.lph` files the unmodified reference wrote (tests/golden/).  numpy + plain loops; small inputs only.

Follows, in /root/reference:
  src/partitioned_mphf.cpp:92-106     re-key of the triplets by minimizer_order, sorted by it
  src/partitioned_mphf.cpp:163-268    mphf::build_inverted_index
  src/quartet_wtree.cpp:13-53         wavelet-tree builder
  include/rs_bit_vector.hpp:120-156   rank directory (no select hints)
  include/ef_sequence.hpp:9-31, 36-75 cumulative_iterator, ef_sequence::encode
  external/pthash/include/encoders/darray.hpp:13-48, 98-120   darray1
  external/pthash/include/encoders/compact_vector.hpp:88-100, 127-145   low bits
"""
import struct

import numpy as np

LEFT, RIGHT_OR_COLLISION, MAXIMAL, NONE = 0, 1, 2, 3  # include/quartet_wtree.hpp:7


def _u64(v):
    return struct.pack("<Q", int(v))


def _vec(words, fmt="<u8"):
    a = np.array(words, dtype=fmt)  # dtype given up front: a list of ints >= 2^63 must not pass through float64
    return _u64(len(a)) + a.tobytes()


def _bits_to_words(bits):
    """bit i of the vector = bit (i % 64) of word i // 64 (bit_vector_builder::push_back)"""
    n = len(bits)
    nwords = (n + 63) // 64
    padded = np.zeros(nwords * 64, dtype=np.uint8)
    padded[:n] = bits
    return np.packbits(padded.reshape(nwords, 64), axis=1, bitorder="little").view("<u8").reshape(nwords)


def rs_bit_vector_image(bits):
    """include/rs_bit_vector.hpp:91-96 (visit) + :120-156 (build_indices, with_select_hints = false)"""
    words = _bits_to_words(np.asarray(bits, dtype=np.uint8))
    pairs = [0]
    next_rank = cur_sub = sub = 0
    for i, w in enumerate(words.tolist()):
        pop = bin(w).count("1")
        shift = i % 8
        if shift:
            sub = ((sub << 9) | cur_sub) & 0xFFFFFFFFFFFFFFFF
        next_rank += pop
        cur_sub += pop
        if shift == 7:
            pairs += [sub, next_rank]
            sub = cur_sub = 0
    for _ in range(8 - len(words) % 8):
        sub = ((sub << 9) | cur_sub) & 0xFFFFFFFFFFFFFFFF
    pairs.append(sub)
    if len(words) % 8:
        pairs += [next_rank, 0]
    return _u64(len(bits)) + _vec(words) + _vec(pairs) + _vec([])


def darray1_image(ones):
    """positions of the set bits -> pthash::darray1 (darray.hpp:13-48 build, :98-120 flush_cur_block)"""
    block_inv, sub_inv, overflow = [], [], []
    for b in range(0, len(ones), 1024):
        cur = ones[b:b + 1024]
        if cur[-1] - cur[0] < (1 << 16):
            block_inv.append(cur[0])
            sub_inv += [cur[i] - cur[0] for i in range(0, len(cur), 32)]
        else:
            block_inv.append(-len(overflow) - 1)
            overflow += cur
            sub_inv += [0xFFFF] * len(range(0, len(cur), 32))
    return _u64(len(ones)) + _vec(block_inv, "<i8") + _vec(sub_inv, "<u2") + _vec(overflow)


def ef_sequence_image(values, universe):
    """ef_sequence::encode(begin, n, u) with the leading zero it adds (include/ef_sequence.hpp:36-75) + visit (:107-112)"""
    n = len(values)
    if n == 0:
        return _u64(0) + _vec([]) + _u64(0) + _vec([], "<i8") + _vec([], "<u2") + _vec([]) + _u64(0) * 3 + _vec([])
    n_enc = n + 1
    q = universe // n_enc
    l = q.bit_length() - 1 if q else 0
    seq = [0] + [int(v) for v in values]
    ones = [(v >> l) + i for i, v in enumerate(seq)]
    high = np.zeros(n_enc + (universe >> l) + 1, dtype=np.uint8)
    high[ones] = 1
    low_words = [0] * ((n_enc * l + 63) // 64 + 1)
    if l:
        mask = (1 << l) - 1
        for i, v in enumerate(seq):
            at = i * l
            low_words[at >> 6] |= ((v & mask) << (at & 63)) & 0xFFFFFFFFFFFFFFFF
            if (at & 63) + l > 64:
                low_words[(at >> 6) + 1] |= (v & mask) >> (64 - (at & 63))
    out = _u64(len(high)) + _vec(_bits_to_words(high))
    out += darray1_image(ones)
    out += _u64(n_enc) + _u64(l) + _u64((1 << l) - 1) + _vec(low_words)
    return out


def classify_minimizers(triplets, orders, k, m):
    """The loop of mphf::build_inverted_index (src/partitioned_mphf.cpp:183-215) over the triplets in minimizer_order
    order.  Returns (types in that order, left_positions, right_or_collision_sizes, none_sizes, none_positions)."""
    by_order = np.argsort(np.asarray(orders), kind="stable")  # the external sort of partitioned_mphf.cpp:93-100
    kinds, left_positions, rc_sizes, none_sizes, none_positions = [], [], [], [], []
    p1s, sizes = triplets["p1"].tolist(), triplets["size"].tolist()
    for t in by_order.tolist():
        p1, size = p1s[t], sizes[t]
        if size == 0:
            kind = RIGHT_OR_COLLISION
            rc_sizes.append(0)
        elif p1 == k - m:
            if size == k - m + 1:
                kind = MAXIMAL
            else:
                kind = RIGHT_OR_COLLISION
                rc_sizes.append(size)
        elif p1 == size - 1:
            kind = LEFT
            left_positions.append(p1 + 1)
        else:
            kind = NONE
            none_positions.append(p1)
            none_sizes.append(size)
        kinds.append(kind)
    return kinds, left_positions, rc_sizes, none_sizes, none_positions


def value_lists(triplets, orders, k, m):
    """the concatenation append_iterator walks (src/partitioned_mphf.cpp:253-265), before the prefix sums"""
    _, lp, rc, nsz, npos = classify_minimizers(triplets, orders, k, m)
    return np.array(lp + rc + nsz + npos, dtype=np.uint64)


def build_inverted_index(triplets, orders, k, m):
    """triplets: records with fields itself, p1, size; orders[i] = minimizer_order(triplets[i].itself).
    Returns ((n_maximal, right_coll_sizes_start, none_sizes_start, none_pos_start), image of wtree + image of
    sizes_and_positions)."""
    kinds, lp, rc, nsz, npos = classify_minimizers(triplets, orders, k, m)
    root = [kd >> 1 for kd in kinds]
    left_right = [kd & 1 for kd in kinds if not kd >> 1]
    max_none = [kd & 1 for kd in kinds if kd >> 1]
    values = np.cumsum(np.array(lp + rc + nsz + npos, dtype=np.uint64))
    universe = int(values[-1]) if len(values) else 0
    body = rs_bit_vector_image(root) + rs_bit_vector_image(left_right) + rs_bit_vector_image(max_none)
    body += ef_sequence_image(values.tolist(), universe)
    rs = len(lp)
    ns = rs + len(rc)
    return (kinds.count(MAXIMAL), rs, ns, ns + len(nsz)), body


def build_inverted_index_alt(triplets, orders):
    """build-u (lphash::mphf_alt): positions = Elias-Fano of the prefix sums of p1, sizes = of the sizes, both in
    minimizer_order order (src/unpartitioned_mphf.cpp:78-96, 152-169).  Returns (num_kmers_in_main_index, image of
    positions + image of sizes)."""
    by_order = np.argsort(np.asarray(orders), kind="stable")
    p1 = np.cumsum(triplets["p1"][by_order].astype(np.uint64))
    size = np.cumsum(triplets["size"][by_order].astype(np.uint64))
    body = ef_sequence_image(p1.tolist(), int(p1[-1]) if len(p1) else 0)
    body += ef_sequence_image(size.tolist(), int(size[-1]) if len(size) else 0)
    return (int(size[-1]) if len(size) else 0), body
