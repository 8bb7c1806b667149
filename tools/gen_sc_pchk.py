#!/usr/bin/env python3
"""Seeded generator of a terminated spatially-coupled (SC-LDPC) parity-check matrix (.pchk) for the sliding-window
decoder (Run_SW_Decoder, dec.cpp:2092-2196). The reference ships no SC-LDPC code, so the window
decoder needs one.

Construction: (3,6)-regular protograph ensemble coupled over `w` = 3 check positions. Position t (0..L-1) holds 2*Z
variable nodes (two protograph types, Z copies each); check position q (0..L+w-2) holds Z check nodes. Variable
(t, type, z) has one edge to check position t+k for k = 0..w-1, through a seeded Z x Z permutation per (t, type, k),
so a check of position q has 2 edges from each of the variable positions q, q-1, q-2 that exist (degree 6 inside,
2 or 4 at the two ends = the rate loss of the termination).
Ordering = what the window decoder assumes: columns by position (t*2Z + type*Z + z), rows by check position (q*Z + c).
Mv[t] = 2Z for t < L (0 beyond), Mc[q] = Z: the per-position node counts Run_SW_Decoder takes (two-sided
termination, code_type 0: SC_D = L + w - 1).

usage: gen_sc_pchk.py Z L SEED OUT.pchk
"""
import sys

import numpy as np

from gen_regular_pchk import _Rng, write_pchk

W = 3


def gen_sc(Z, L, seed):
    """-> (M, N, row_ptr, col_idx, Mv[D], Mc[D])"""
    rng = _Rng(seed)
    D = L + W - 1
    N, M = 2 * Z * L, Z * D
    per_row = [[] for _ in range(M)]
    for t in range(L):
        for typ in range(2):
            for k in range(W):
                perm = rng.permutation(Z)
                for z in range(Z):
                    per_row[(t + k) * Z + perm[z]].append(t * 2 * Z + typ * Z + z)
    row_ptr = np.zeros(M + 1, dtype=np.int32)
    col_idx = []
    for i in range(M):
        per_row[i].sort()
        col_idx.extend(per_row[i])
        row_ptr[i + 1] = len(col_idx)
    Mv = np.array([2 * Z if t < L else 0 for t in range(D)], dtype=np.int32)
    Mc = np.full(D, Z, dtype=np.int32)
    return M, N, row_ptr, np.asarray(col_idx, dtype=np.int32), Mv, Mc


def main(argv):
    if len(argv) != 5:
        sys.stderr.write(__doc__)
        return 2
    Z, L, seed = (int(x) for x in argv[1:4])
    M, N, row_ptr, col_idx, Mv, Mc = gen_sc(Z, L, seed)
    write_pchk(argv[4], M, N, row_ptr, col_idx)
    print("wrote %s: M=%d N=%d E=%d L=%d w=%d Mv=%d Mc=%d" % (argv[4], M, N, len(col_idx), L, W, 2 * Z, Z))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
