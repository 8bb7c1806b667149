#!/usr/bin/env python3
"""alist <-> .pchk conversion (replaces the reference's alist-to-pchk.cpp, which is not in the shipped exe).

The alist dialect is the one alist-to-pchk.cpp:76-134 reads: `M N`, `max_row_weight max_col_weight`, M row weights,
N column weights, then M row lists (1-based column indices, 0-padded to max_row_weight) and N column lists (1-based row
indices, 0-padded); the two lists must describe the same matrix. `-t` transposes (alist-to-pchk.cpp:45-60).
.pchk = rcode.cpp:54-86 / mod2sparse.cpp:338-376.

usage: alist_to_pchk.py [-t] ALIST PCHK        |   alist_to_pchk.py --to-alist PCHK ALIST
"""
import struct
import sys


class AlistError(ValueError):
    pass


def read_alist(path):
    """-> (M, N, rows) with rows[i] = sorted list of 0-based columns. Raises AlistError like bad_alist_file()."""
    try:
        tok = [int(t) for t in open(path).read().split()]
    except ValueError:
        raise AlistError("Alist file doesn't have the right format")
    pos = 0

    def nxt():
        nonlocal pos
        if pos >= len(tok):
            raise AlistError("Alist file doesn't have the right format")
        pos += 1
        return tok[pos - 1]
    M, N = nxt(), nxt()
    if M < 1 or N < 1:
        raise AlistError("Alist file doesn't have the right format")
    mxrw, mxcw = nxt(), nxt()
    if not (0 <= mxrw <= N and 0 <= mxcw <= M):
        raise AlistError("Alist file doesn't have the right format")
    rw = [nxt() for _ in range(M)]
    cw = [nxt() for _ in range(N)]
    if any(not 0 <= w <= N for w in rw) or any(not 0 <= w <= M for w in cw):
        raise AlistError("Alist file doesn't have the right format")
    rows = [set() for _ in range(M)]
    tot = 0
    for i in range(M):
        for k in range(mxrw):
            j = nxt()
            if j < 0 or j > N or (k >= rw[i] and j != 0) or (k < rw[i] and j == 0):
                raise AlistError("Alist file doesn't have the right format")
            if j == 0:
                continue
            if j - 1 in rows[i]:
                raise AlistError("Alist file doesn't have the right format")
            rows[i].add(j - 1)
            tot += 1
    for j in range(N):
        for k in range(mxcw):
            i = nxt()
            if i < 0 or i > M or (k >= cw[j] and i != 0) or (k < cw[j] and i == 0):
                raise AlistError("Alist file doesn't have the right format")
            if i == 0:
                continue
            if j not in rows[i - 1]:
                raise AlistError("Alist file doesn't have the right format")
            tot -= 1
    if tot != 0 or pos != len(tok):
        raise AlistError("Alist file doesn't have the right format")
    return M, N, [sorted(r) for r in rows]


def transpose(M, N, rows):
    cols = [[] for _ in range(N)]
    for i, r in enumerate(rows):
        for j in r:
            cols[j].append(i)
    return N, M, cols


def write_pchk(path, M, N, rows):
    out = [struct.pack("<iii", (ord("P") << 8) + 0x80, M, N)]
    for i, r in enumerate(rows):
        if r:
            out.append(struct.pack("<%di" % (len(r) + 1), -(i + 1), *[j + 1 for j in r]))
    out.append(struct.pack("<i", 0))
    open(path, "wb").write(b"".join(out))


def read_pchk(path):
    data = open(path, "rb").read()
    w = struct.unpack("<%di" % (len(data) // 4), data[:len(data) // 4 * 4])
    if not w or w[0] != (ord("P") << 8) + 0x80:
        raise ValueError("File %s doesn't contain a parity check matrix" % path)
    M, N = w[1], w[2]
    rows = [set() for _ in range(M)]
    row = -1
    for v in w[3:]:
        if v == 0:
            break
        if v < 0:
            row = -v - 1
        else:
            rows[row].add(v - 1)
    return M, N, [sorted(r) for r in rows]


def write_alist(path, M, N, rows):
    _, _, cols = transpose(M, N, rows)
    mxrw = max((len(r) for r in rows), default=0)
    mxcw = max((len(c) for c in cols), default=0)
    L = ["%d %d" % (M, N), "%d %d" % (mxrw, mxcw), " ".join(str(len(r)) for r in rows), " ".join(str(len(c)) for c in cols)]
    for r in rows:
        L.append(" ".join(str(j + 1) for j in r) + " 0" * (mxrw - len(r)))
    for c in cols:
        L.append(" ".join(str(i + 1) for i in c) + " 0" * (mxcw - len(c)))
    open(path, "w").write("\n".join(L) + "\n")


def main(argv):
    a = argv[1:]
    if len(a) == 3 and a[0] == "--to-alist":
        write_alist(a[2], *read_pchk(a[1]))
        return 0
    trans = False
    while a and a[0] == "-t":
        trans, a = True, a[1:]
    if len(a) != 2:
        sys.stderr.write("Usage: alist-to-pchk [ -t ] alist-file pchk-file\n")
        return 1
    try:
        M, N, rows = read_alist(a[0])
    except OSError:
        sys.stderr.write("Can't open alist file: %s\n" % a[0])
        return 1
    except AlistError as e:
        sys.stderr.write(str(e) + "\n")
        return 1
    if trans:
        M, N, rows = transpose(M, N, rows)
    write_pchk(a[1], M, N, rows)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
