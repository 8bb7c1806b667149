#!/usr/bin/env python3
"""Seeded generator of Neal-style random column-regular LDPC parity-check matrices (.pchk).

The reference ships no code generator (Neal's make-ldpc is absent), so BASELINE config 5
(N=65536, column weight 3, rate 0.9) needs one. Construction ("evenboth"-like): `wc` layers; in layer t a seeded
permutation spreads the N columns as evenly as possible over the M rows; columns with a repeated row or that
close a 4-cycle with an earlier column are repaired by swapping with a random other column of the layer.

Output format = the reference's .pchk (rcode.cpp:54-86 / mod2sparse.cpp:338-376): int32 LE 0x5080, M, N,
then per row -(i+1) followed by col+1 ascending, terminator 0.

usage: gen_regular_pchk.py N M WC SEED OUT.pchk
"""
import struct
import sys

import numpy as np

_M64 = (1 << 64) - 1


def _mix(z):
    z &= _M64
    z ^= z >> 30; z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27; z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z


class _Rng:
    """splitmix64 stream - deterministic across numpy versions."""

    def __init__(self, seed):
        self.s = _mix(seed * 0x9E3779B97F4A7C15 + 0x1234567)

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        return _mix(self.s)

    def below(self, n):
        return self.next() % n

    def permutation(self, n):
        p = list(range(n))
        for i in range(n - 1, 0, -1):
            j = self.below(i + 1)
            p[i], p[j] = p[j], p[i]
        return p


def gen_regular(N, M, wc, seed, no4cycle=True, max_pass=200):
    """-> (row_ptr[M+1], col_idx[E]) int32 CSR, rows sorted by column."""
    rng = _Rng(seed)
    rows = [[0] * N for _ in range(wc)]
    for t in range(wc):
        perm = rng.permutation(N)
        for j in range(N):
            rows[t][j] = perm[j] % M
    pair_owner = {}

    def col_pairs(j):
        r = sorted(rows[t][j] for t in range(wc))
        return [(r[a], r[b]) for a in range(wc) for b in range(a + 1, wc)]

    def conflicts(j):
        r = [rows[t][j] for t in range(wc)]
        c = wc - len(set(r))
        if no4cycle:
            for p in col_pairs(j):
                o = pair_owner.get(p)
                if o is not None and o != j:
                    c += 1
        return c

    def claim(j):
        for p in col_pairs(j):
            pair_owner[p] = j

    def release(j):
        for p in col_pairs(j):
            if pair_owner.get(p) == j:
                del pair_owner[p]

    pending = []
    for j in range(N):
        if conflicts(j):
            pending.append(j)
        else:
            claim(j)
    # Repair: swap one layer's row of a conflicting column j with that of a healthy column k whenever k stays
    # healthy and j's conflict count does not grow (sideways moves allowed: a column can conflict in all wc pairs).
    pset = set(pending)
    for j in pending:
        cj = conflicts(j)
        tries = 0
        while cj:
            tries += 1
            if tries > max_pass * 64:
                raise RuntimeError("could not repair column %d" % j)
            t = rng.below(wc)
            k = rng.below(N)
            if k == j or k in pset:
                continue
            release(k)
            rows[t][j], rows[t][k] = rows[t][k], rows[t][j]
            nj = conflicts(j)
            if conflicts(k) == 0 and nj <= cj and not (set(col_pairs(j)) & set(col_pairs(k))):
                cj = nj
            else:
                rows[t][j], rows[t][k] = rows[t][k], rows[t][j]
            claim(k)
        claim(j)
        pset.discard(j)
    per_row = [[] for _ in range(M)]
    for j in range(N):
        for t in range(wc):
            per_row[rows[t][j]].append(j)
    row_ptr = np.zeros(M + 1, dtype=np.int32)
    col_idx = []
    for i in range(M):
        per_row[i].sort()
        col_idx.extend(per_row[i])
        row_ptr[i + 1] = len(col_idx)
    return row_ptr, np.asarray(col_idx, dtype=np.int32)


def write_pchk(path, M, N, row_ptr, col_idx):
    out = [struct.pack("<iii", (ord("P") << 8) + 0x80, M, N)]
    for i in range(M):
        a, b = int(row_ptr[i]), int(row_ptr[i + 1])
        if a == b:
            continue
        out.append(struct.pack("<i", -(i + 1)))
        out.append((np.asarray(col_idx[a:b], dtype="<i4") + 1).tobytes())
    out.append(struct.pack("<i", 0))
    with open(path, "wb") as f:
        f.write(b"".join(out))


def main(argv):
    if len(argv) != 6:
        sys.stderr.write(__doc__)
        return 2
    N, M, wc, seed = (int(x) for x in argv[1:5])
    row_ptr, col_idx = gen_regular(N, M, wc, seed)
    write_pchk(argv[5], M, N, row_ptr, col_idx)
    print("wrote %s: M=%d N=%d E=%d" % (argv[5], M, N, len(col_idx)))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
