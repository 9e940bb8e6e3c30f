#!/usr/bin/env python
"""Generator + verifier of fsq_median_pair.cuh: the 5x5 medians of TWO vertically adjacent windows at once.

The windows of output rows y and y+1 share four of their five rows.  Of the 20 shared values only the middle six
(ranks 7..12) can be the median of either window (the 7 smallest and the 7 largest have >= 13 values of the window on
the other side), so:
   sort each of the six 5-pixel rows (9 compare-exchanges each, the optimal 5-sorter),
   merge the four shared rows (odd-even merges 5+5, 5+5, 10+10) keeping only what ranks 7..12 need,
   median(window) = rank 6 of (six sorted shared values + five sorted own values) = min_i max(M[i-1], O[5-i]).
Every operation is a min or a max, so the 0-1 principle applies: the generated program is checked on all 2^25 0/1
inputs of each window (`--verify`), and dead operations are removed by a liveness pass before emission.

    python tools/gen_median_pair.py [--verify] > fluorosequencingimageanalysis_b200/csrc/fsq_median_pair.cuh
"""
import sys

SORT5 = [(0, 1), (3, 4), (2, 4), (2, 3), (0, 3), (0, 2), (1, 4), (1, 3), (1, 2)]


class Prog(object):
    def __init__(self, n_in):
        self.ops = []            # (kind, dst, a, b) with kind in {'min', 'max'}; values are SSA ids
        self.n = n_in

    def new(self, kind, a, b):
        self.ops.append((kind, self.n, a, b))
        self.n += 1
        return self.n - 1

    def ce(self, w, i, j):       # compare-exchange on a wire list: w[i] <- min, w[j] <- max
        a, b = w[i], w[j]
        w[i], w[j] = self.new('min', a, b), self.new('max', a, b)


def oddeven_merge(P, w, ia, ib):
    """Batcher's odd-even merge of the sorted wires w[ia] and w[ib] (index lists, any lengths); on return the
    concatenation ia + ib... is NOT in place: returns the list of wire indices in sorted order."""
    if not ia:
        return list(ib)
    if not ib:
        return list(ia)
    if len(ia) == 1 and len(ib) == 1:
        P.ce(w, ia[0], ib[0])
        return [ia[0], ib[0]]
    ev = oddeven_merge(P, w, ia[0::2], ib[0::2])      # merged even-indexed elements (0-based 0, 2, ...)
    od = oddeven_merge(P, w, ia[1::2], ib[1::2])
    # interleave: result = ev[0], then pairs (od[k], ev[k+1]) compare-exchanged
    out = [ev[0]]
    k = 0
    while k < len(od) and k + 1 < len(ev):
        P.ce(w, od[k], ev[k + 1])
        out += [od[k], ev[k + 1]]
        k += 1
    out += od[k:] + ev[k + 1:]
    return out


def build(presorted=False):
    P = Prog(30)
    w = list(range(30))                                  # wire r*5+c holds element c of row r (rows 0..5)
    for r in range(6):
        if presorted:
            continue                                     # the caller hands in rows that are already sorted
        for (i, j) in SORT5:
            P.ce(w, r * 5 + i, r * 5 + j)
    rows = [[r * 5 + c for c in range(5)] for r in range(6)]
    m01 = oddeven_merge(P, w, rows[1], rows[2])
    m23 = oddeven_merge(P, w, rows[3], rows[4])
    m = oddeven_merge(P, w, m01, m23)                     # 20 sorted shared values
    mid = [w[i] for i in m[7:13]]
    outs = []
    for own in (rows[0], rows[5]):
        O = [w[i] for i in own]
        terms = [P.new('max', mid[i], O[4 - i]) for i in range(5)] + [mid[5]]
        acc = terms[0]
        for t in terms[1:]:
            acc = P.new('min', acc, t)
        outs.append(acc)
    return P, outs


def prune(P, outs):
    live = set(outs)
    keep = []
    for op in reversed(P.ops):
        if op[1] in live:
            keep.append(op)
            live.add(op[2]); live.add(op[3])
    return keep[::-1]


def verify(ops, outs, presorted=False):
    import numpy as np
    chunk = 1 << 21
    for which, rows in ((0, range(0, 5)), (1, range(1, 6))):
        for base in range(0, 1 << 25, chunk):
            v = np.arange(base, base + chunk, dtype=np.uint32)
            val = {}
            ones = np.zeros(chunk, dtype=np.uint8)
            for r in range(6):
                for c in range(5):
                    if r in rows:
                        bit = (list(rows).index(r)) * 5 + c
                        x = ((v >> bit) & 1).astype(np.uint8)
                        ones += x
                    else:
                        x = ((v >> ((r * 7 + c) % 25)) & 1).astype(np.uint8)       # arbitrary: must not matter
                    val[r * 5 + c] = x
            if presorted:                                # sort every row first, as the kernel's pre-pass does
                for r in range(6):
                    for (i, j) in SORT5:
                        a, b = val[r * 5 + i], val[r * 5 + j]
                        val[r * 5 + i], val[r * 5 + j] = a & b, a | b
            for kind, d, a, b in ops:
                val[d] = (val[a] & val[b]) if kind == 'min' else (val[a] | val[b])
            assert np.array_equal(val[outs[which]], (ones >= 13).astype(np.uint8)), (which, base)
    return True


def emit_sorted(ops, outs):
    """median25_pair_sorted: the same selection for rows that are already sorted (shared-memory pre-pass), + sort5."""
    o = []
    o.append("// (continued) rows pre-sorted by sort5(): %d min/max operations for both windows." % len(ops))
    o.append("namespace fsq {")
    o.append("template <typename V>")
    o.append("__device__ __forceinline__ void sort5(V (&r)[5]) {")
    for (i, j) in SORT5:
        o.append("    { const V lo = min(r[%d], r[%d]), hi = max(r[%d], r[%d]); r[%d] = lo; r[%d] = hi; }" % (i, j, i, j, i, j))
    o.append("}")
    o.append("template <typename V>")
    o.append("__device__ __forceinline__ void median25_pair_sorted(const V (&p)[30], V& top, V& bottom) {")
    name = lambda i: ("p[%d]" % i) if i < 30 else ("t%d" % i)
    for kind, d, a, b in ops:
        o.append("    const V t%d = %s(%s, %s);" % (d, kind, name(a), name(b)))
    o.append("    top = %s; bottom = %s;" % (name(outs[0]), name(outs[1])))
    o.append("}")
    o.append("}  // namespace fsq")
    return "\n".join(o) + "\n"


def emit(ops, outs):
    o = []
    o.append("// GENERATED by tools/gen_median_pair.py (verified on all 2^25 0/1 inputs of each window) -- do not edit.")
    o.append("// Medians of the two vertically adjacent 5x5 windows formed by six rows of five values: top = rows 0..4,")
    o.append("// bottom = rows 1..5.  %d min/max operations for both (the 99-comparator network: 198 per window)." % len(ops))
    o.append("#pragma once")
    o.append("namespace fsq {")
    o.append("template <typename V>")
    o.append("__device__ __forceinline__ void median25_pair(const V (&p)[30], V& top, V& bottom) {")
    name = lambda i: ("p[%d]" % i) if i < 30 else ("t%d" % i)
    for kind, d, a, b in ops:
        o.append("    const V t%d = %s(%s, %s);" % (d, kind, name(a), name(b)))
    o.append("    top = %s; bottom = %s;" % (name(outs[0]), name(outs[1])))
    o.append("}")
    o.append("}  // namespace fsq")
    return "\n".join(o) + "\n"


def generate(check=False):
    P, outs = build()
    ops = prune(P, outs)
    P2, outs2 = build(presorted=True)
    ops2 = prune(P2, outs2)
    if check:
        verify(ops, outs)
        verify(ops2, outs2, presorted=True)
    return emit(ops, outs) + emit_sorted(ops2, outs2), len(ops), len(ops2)


if __name__ == "__main__":
    text, n1, n2 = generate("--verify" in sys.argv)
    if "--verify" in sys.argv:
        sys.stderr.write("verified: %d operations (raw rows), %d (pre-sorted rows)\n" % (n1, n2))
    sys.stdout.write(text)
