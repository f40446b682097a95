"""Search for small LOP3 circuits of K1's per-plane truth tables (hawk_core.h, pack_chunk_v3).

A plane word of K1 is a boolean function of the letter number's five bits (b0..b4 = ASCII bits
0..4), evaluated for 32 characters at once. Only 16 of the 32 letter numbers matter: the 15 IUPAC
letters and 0 (NUL, the unused slot, which must give zero planes) -- any other byte makes the
batch invalid as a whole, so the planes are don't-care there, while the `valid` function must be
exact on all 32. This script finds, per function, the smallest circuit of 3-input gates of the form

    f = L3(L1(x, y, z), L2(u, v, w), t)          (3 gates)     or     f = L2(L1(x, y, z), u, v)   (2 gates)

by exhaustive search over the first-level gates and prints the LOP3 immediates.

    python tools/lop3_search.py
"""

import itertools

import numpy as np

LETTERS = "ACGTRYSWKMBDHVN"
CODES = [1, 2, 4, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15]
IDX = [ord(c) & 31 for c in LETTERS]
FULL = 0xFFFFFFFF
# truth tables of the inputs over the 32 letter numbers: bit L of B[k] = bit k of L
B = [sum(((L >> k) & 1) << L for L in range(32)) for k in range(5)]


def table_of(bit):
    return sum(1 << i for i, c in zip(IDX, CODES) if c & (1 << bit))


CARE = sum(1 << i for i in IDX) | 1  # the letters and NUL
VALID = sum(1 << i for i in IDX)


def lop3(imm, a, b, c):
    r = 0
    for i in range(8):
        if imm & (1 << i):
            r |= (a if i & 4 else ~a) & (b if i & 2 else ~b) & (c if i & 1 else ~c)
    return r & FULL


def first_level():
    """all distinct functions L(x, y, z) of three of the five inputs -> {table: (imm, (x, y, z))}"""
    out = {}
    for tri in itertools.combinations(range(5), 3):
        for imm in range(256):
            t = lop3(imm, B[tri[0]], B[tri[1]], B[tri[2]])
            out.setdefault(t, (imm, tri))
    return out


def outer_imm(f, care, ins):
    """immediate of L(ins[0], ins[1], ins[2]) equal to f on `care`, or None if f is not a function of them"""
    imm = 0
    for i in range(8):
        m = care
        for k, t in enumerate(ins):
            m &= t if i & (4 >> k) else ~t
        m &= FULL
        if m & f and m & ~f:
            return None
        if m & f:
            imm |= 1 << i
    return imm


def search(f, care, name):
    fl = first_level()
    tabs = list(fl)
    # 1 gate
    for t in tabs:
        if (t ^ f) & care == 0:
            print(f"{name}: 1 gate  L{fl[t]}")
            return
    # 2 gates: f = L(g1, bu, bv)
    for t in tabs:
        for u, v in itertools.combinations(range(5), 2):
            imm = outer_imm(f, care, (t, B[u], B[v]))
            if imm is not None:
                print(f"{name}: 2 gates g1 = L{fl[t]}; f = L(imm={imm:#04x}; g1, b{u}, b{v})")
                return
    # 3 gates: f = L(g1, g2, bt) -- vectorised over g2
    arr = np.array(tabs, dtype=np.uint64)
    f64, c64 = np.uint64(f), np.uint64(care)
    for t1 in tabs:
        for bt in range(5):
            ok = np.ones(len(arr), bool)
            for i in range(8):
                m = np.full(len(arr), c64, np.uint64)
                m &= np.uint64(t1 if i & 4 else ~t1 & FULL)
                m &= arr if i & 2 else ~arr & np.uint64(FULL)
                m &= np.uint64(B[bt] if i & 1 else ~B[bt] & FULL)
                ok &= ~(((m & f64) != 0) & ((m & ~f64 & np.uint64(FULL)) != 0))
            hit = np.flatnonzero(ok)
            if len(hit):
                t2 = tabs[int(hit[0])]
                imm = outer_imm(f, care, (t1, t2, B[bt]))
                print(f"{name}: 3 gates g1 = L{fl[t1]}; g2 = L{fl[t2]}; f = L(imm={imm:#04x}; g1, g2, b{bt})")
                return
    print(f"{name}: no circuit of <= 3 gates in this family")


if __name__ == "__main__":
    for bit, nm in enumerate("ACGT"):
        search(table_of(bit), CARE, f"plane {nm} (don't-care off the letters)")
    search(VALID, FULL, "valid (exact)")
