"""CPU tier: a pure-Python model of csrc/witness_fast.cuh -- the same transform sizes, wrap-around fix-ups, leaf Horner steps,
tree levels, short trailing blocks and the index arithmetic of wf_pass -- against the C oracle's restatement of the
reference (interpolate / multiply / divide, util/polynomials.tcc:9-81).  It pins the ALGORITHM and the host-side table
layout (what ensure_fast_tables builds in rsgpu.cu) without a GPU; the kernels themselves are compared with the dense path,
the oracle and the reference in tests/test_gpu_parity.py."""
import random

import numpy as np
import pytest

import oracle_lib as O

p = 786433  # 3 * 2^18 + 1: has 2^13-th roots of unity for the model's transforms
def inv(a): return pow(a, p - 2, p)
def prim_root_2n(n2):  # minimal primitive n2-th root like SEAL (any primitive root works for the model)
    for g in range(2, 1000):
        r = pow(g, (p - 1) // n2, p)
        if pow(r, n2 // 2, p) == p - 1: return r
def bitrev(x, bits):
    return int(format(x, '0%db' % bits)[::-1], 2) if bits else 0
LOGN = 12; N = 1 << LOGN
psi = prim_root_2n(2 * N)
fwd = [1] * N; invt = [1] * N
for i in range(1, N):
    k = bitrev(i, LOGN); fwd[k] = pow(psi, i, p); invt[k] = inv(fwd[k])

def h_ntt_fwd(a, lg):
    n = 1 << lg; a = a[:]
    for s in range(lg):
        gap = n >> (s + 1)
        for blk in range(1 << s):
            w = fwd[(1 << s) + blk]
            for o in range(gap):
                i = blk * 2 * gap + o
                x, y = a[i], a[i + gap] * w % p
                a[i], a[i + gap] = (x + y) % p, (x - y) % p
    return a

def wf_pass(buf, nb, lg, s, RL, inverse):
    R = 1 << RL; lgi = lg - RL; lgg = lg - s - RL; g = 1 << lgg
    for r in range(nb << lgi):
        b = r >> lgi; li = r & ((1 << lgi) - 1); o = li & (g - 1); blk = li >> lgg
        base = (b << lg) + (blk << (lg - s)) + o
        v = [buf[base + k * g] for k in range(R)]
        us = range(RL) if not inverse else range(RL - 1, -1, -1)
        for u in us:
            half = R >> (u + 1); tbase = (1 << (s + u)) + (blk << u)
            for grp in range(1 << u):
                for k in range(half):
                    i0 = grp * 2 * half + k; i1 = i0 + half
                    if not inverse:
                        w = fwd[tbase + grp]; x, y = v[i0], v[i1] * w % p
                        v[i0], v[i1] = (x + y) % p, (x - y) % p
                    else:
                        w = invt[tbase + grp]; x, y = v[i0], v[i1]
                        v[i0], v[i1] = (x + y) % p, (x - y) * w % p
        for k in range(R): buf[base + k * g] = v[k]
def ntt_fwd(buf, nb, lg, s0):
    s = s0
    while lg - s >= 4: wf_pass(buf, nb, lg, s, 4, False); s += 4
    if lg - s: wf_pass(buf, nb, lg, s, lg - s, False)
def ntt_inv(buf, nb, lg):
    rem = lg; first = rem & 3
    if first: wf_pass(buf, nb, lg, rem - first, first, True)
    rem -= first
    while rem: wf_pass(buf, nb, lg, rem - 4, 4, True); rem -= 4

B = 16; WC_MAX = 32; HMAX = 16
def shape(n):
    lg = 5
    while (1 << lg) < n: lg += 1
    s = 1 << lg
    if 2 * n - 1 > s + WC_MAX: lg += 1; s <<= 1
    return s, lg, max(0, 2 * n - 1 - s)
def polymul(a, b):
    r = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b): r[i + j] = (r[i + j] + x * y) % p
    return r
def wrapped(u, lu, v, lv, S, wc):
    out = []
    for k in range(wc):
        deg = k + S; acc = 0
        i = deg - lv + 1 if deg >= lv else 0
        while i < lu and i <= deg: acc += u[i] * v[deg - i]; i += 1
        out.append(acc % p)
    return out

def tables(n):
    S, logS, wc = shape(n)
    fact = [1] * (n + 1)
    for i in range(1, n + 1): fact[i] = fact[i - 1] * i % p
    ifact = [inv(f) for f in fact]
    g = [(p - ifact[i]) % p if i & 1 else ifact[i] for i in range(n)]
    invS = inv(S)
    Ghat = [x * invS % p for x in h_ntt_fwd(g + [0] * (S - n), logS)]
    npad = (n + B - 1) // B * B
    tree = []
    for r in range(npad // B):
        f = [1]
        for x in range(r * B, (r + 1) * B): f = polymul(f, [(-x) % p, 1])
        tree.append(f)
    Phat, Pnat = [], []
    m = B
    while m < n:
        two_m = 2 * m; nb_active = (n - m + two_m - 1) // two_m; lg = two_m.bit_length() - 1
        ph = [0] * S
        for b in range(nb_active):
            a = tree[2 * b] + [0] * (two_m - m - 1)
            a = h_ntt_fwd(a, lg)
            for i in range(two_m): ph[b * two_m + i] = a[i] * inv(two_m) % p
        Phat.append(ph); Pnat.append(tree[2 * (nb_active - 1)])
        tree = [polymul(tree[2 * r], tree[2 * r + 1]) for r in range(len(tree) // 2)]
        m <<= 1
    # Z and u
    Z = [1]
    for x in range(n): Z = polymul(Z, [(-x) % p, 1])
    lu = n - 1
    u = [0] * max(lu, 1); u[0] = 1
    for i in range(1, lu):
        acc = sum(Z[n - t] * u[i - t] for t in range(1, i + 1)) % p
        u[i] = (-acc) % p
    Vhat = [x * invS % p for x in h_ntt_fwd(u[:lu] + [0] * (S - lu), logS)]
    return dict(S=S, logS=logS, wc=wc, ifact=ifact, g=g, Ghat=Ghat, Phat=Phat, Pnat=Pnat, Z=Z, u=u, Vhat=Vhat, invS=invS)

def interp_fast(y, n, T):
    S, logS, wc = T['S'], T['logS'], T['wc']
    A = [y[i] * T['ifact'][i] % p if i < n else 0 for i in range(S)]
    wr = wrapped(A, n, T['g'], n, S, wc)
    ntt_fwd(A, 1, logS, 0)
    A = [A[i] * T['Ghat'][i] % p for i in range(S)]
    ntt_inv(A, 1, logS)
    A = [((A[i] + (wr[i] if i < wc else 0)) % p if i < n else 0) for i in range(S)]
    # leaves
    npad = (n + B - 1) // B * B
    for blk in range(npad // B):
        c = A[blk * B:(blk + 1) * B]
        for k in range(B - 2, -1, -1):
            pt = blk * B + k
            for j in range(k, B - 1): c[j] = (c[j] - c[j + 1] * pt) % p
        A[blk * B:(blk + 1) * B] = c
    m = B; lvl = 0
    Bf = [0] * S
    while m < n:
        lg = (2 * m).bit_length() - 1; two_m = 2 * m
        nb_active = (n - m + two_m - 1) // two_m; last = nb_active - 1
        h_last = min(m, n - (last * two_m + m)); shortp = h_last <= HMAX; nbN = nb_active - (1 if shortp else 0)
        for b in range(nbN):
            for i in range(m):
                v = A[b * two_m + m + i]; Bf[b * two_m + i] = v; Bf[b * two_m + m + i] = v
        hs = [A[last * two_m + m + i] for i in range(h_last)] if shortp else []
        if nbN:
            ntt_fwd(Bf, nbN, lg, 1)
            for idx in range(nbN * two_m): Bf[idx] = Bf[idx] * T['Phat'][lvl][idx] % p
            ntt_inv(Bf, nbN, lg)
            for idx in range(nbN * two_m):
                x = Bf[idx]
                if (idx & (two_m - 1)) < m: x = (x + A[idx]) % p
                A[idx] = x
        if shortp:
            Pn = T['Pnat'][lvl]
            for j in range(two_m):
                acc = 0; i = j - m if j > m else 0
                while i < h_last and i <= j: acc += hs[i] * Pn[j - i]; i += 1
                x = acc % p
                if j < m: x = (x + A[last * two_m + j]) % p
                A[last * two_m + j] = x
        m <<= 1; lvl += 1
    return A[:n]

def quotient_fast(a, b, n, T):
    S, logS, wc = T['S'], T['logS'], T['wc']
    A = a + [0] * (S - n); Bb = b + [0] * (S - n)
    wr = wrapped(A, n, Bb, n, S, wc)
    ntt_fwd(A, 1, logS, 0); ntt_fwd(Bb, 1, logS, 0)
    A = [x * y % p for x, y in zip(A, Bb)]
    ntt_inv(A, 1, logS)
    lu = n - 1
    U = [0] * S
    for i in range(lu):
        k = 2 * n - 2 - i
        U[i] = wr[k - S] if k >= S else A[k] * T['invS'] % p
    wc2 = 2 * lu - 1 - S if 2 * lu > S + 1 else 0
    wr2 = wrapped(U, lu, T['u'], lu, S, wc2)
    ntt_fwd(U, 1, logS, 0)
    U = [U[i] * T['Vhat'][i] % p for i in range(S)]
    ntt_inv(U, 1, logS)
    H = []
    for i in range(lu):
        k = lu - 1 - i
        H.append((U[k] + (wr2[k] if k < wc2 else 0)) % p)
    return H

def lagrange(y, n):
    # coefficients of the interpolant on 0..n-1
    res = [0] * n
    Z = [1]
    for x in range(n): Z = polymul(Z, [(-x) % p, 1])
    for x in range(n):
        b = [0] * n; b[n - 1] = 1
        for k in range(n - 1, 0, -1): b[k - 1] = (Z[k] + x * b[k]) % p
        d = 1
        for i in range(n):
            if i != x: d = d * (x - i) % p
        s = y[x] * inv(d) % p
        for k in range(n): res[k] = (res[k] + b[k] * s) % p
    return res
def divmod_Z(P, Z, n):
    P = P[:]; q = [0] * (len(P) - n)
    for i in range(len(P) - 1, n - 1, -1):
        c = P[i]; q[i - n] = c
        for k in range(n + 1): P[i - n + k] = (P[i - n + k] - c * Z[k]) % p
    return q



@pytest.mark.parametrize("n", [2, 3, 15, 16, 17, 31, 32, 33, 47, 48, 49, 64, 65, 81, 129, 257, 300])
def test_model_matches_oracle(n):
    random.seed(n)
    T = tables(n)
    y = [random.randrange(p) for _ in range(n)]
    got = interp_fast(y, n, T)
    want = O.interpolate(np.array(y, dtype=np.uint64).reshape(n, 1), 1, 1, [p])[:, 0]
    assert got == [int(x) for x in want]
    a = [random.randrange(p) for _ in range(n)]
    b = [random.randrange(p) for _ in range(n)]
    H = quotient_fast(a, b, n, T)
    col = lambda v: np.array(v, dtype=np.uint64).reshape(n, 1)
    Hw, hl = O.witness_H(col(a), col(b), col([0] * n), 1, 1, [p])
    assert hl == n - 1 and H == [int(x) for x in Hw[:, 0]]


# ---- blocked products (witness_fast.cuh, k_interp_big / k_quotient_big): n beyond what one transform of size TS <= N_E serves ----
# Polynomials are cut into blocks of h = TS/2 coefficients; a block product (< 2h coefficients) fits one negacyclic transform
# of size TS without wrapping, the blocks of a constant are held transformed, and the partial products of one output block are
# summed in the transform domain before its single inverse transform.  n <= 4h = 2*TS.
def blocks_fwd(src, nblk, h, lgT):
    """[nblk][TS] transformed copies of the h-coefficient blocks of src (zero-padded)"""
    TS = 2 * h
    X = [0] * (nblk * TS)
    for j in range(nblk):
        for i in range(h):
            if j * h + i < len(src): X[j * TS + i] = src[j * h + i]
    ntt_fwd(X, nblk, lgT, 0)
    return X
def block_out(X, nx, C, nc, k, TS, lgT, scale=1):
    """inverse transform of sum_{i+j=k, i<nx, j<nc} X_i * C_j (coefficients k*h .. k*h+2h-1 of that part of the product)"""
    T = [0] * TS
    for i in range(nx):
        j = k - i
        if 0 <= j < nc:
            for e in range(TS): T[e] = (T[e] + X[i * TS + e] * C[j * TS + e]) % p
    if scale != 1: T = [x * scale % p for x in T]
    ntt_inv(T, 1, lgT)
    return T
def polymul_big(a, b):   # any product; the model's tables use the schoolbook one (the host code: blocked transforms)
    return polymul(a, b)

def tables_big(n, TS):
    h = TS // 2; lgT = TS.bit_length() - 1
    Sb = 2 * h if n <= 2 * h else 4 * h
    assert h + B <= n <= 4 * h
    nx = (n + h - 1) // h
    fact = [1] * (n + 1)
    for i in range(1, n + 1): fact[i] = fact[i - 1] * i % p
    ifact = [inv(f) for f in fact]
    g = [(p - ifact[i]) % p if i & 1 else ifact[i] for i in range(n)]
    invT = inv(TS)
    Gblk = [x * invT % p for x in blocks_fwd(g, nx, h, lgT)]
    npad = (n + B - 1) // B * B
    tree = []
    for r in range(npad // B):
        f = [1]
        for x in range(r * B, (r + 1) * B): f = polymul(f, [(-x) % p, 1])
        tree.append(f)
    Phat, Pnat, Ptop = [], [], None
    m = B
    while m < n:
        two_m = 2 * m
        if two_m <= TS:
            nb_active = (n - m + two_m - 1) // two_m; lg = two_m.bit_length() - 1
            ph = [0] * Sb
            for b in range(nb_active):
                a = h_ntt_fwd(tree[2 * b] + [0] * (two_m - m - 1), lg)
                for i in range(two_m): ph[b * two_m + i] = a[i] * inv(two_m) % p
            Phat.append(ph); Pnat.append(tree[2 * (nb_active - 1)])
        else:     # m = TS: P = x^m + (two blocks), the blocks held transformed
            assert m == TS and len(tree[0]) == m + 1
            Ptop = [x * invT % p for x in blocks_fwd(tree[0][:m], 2, h, lgT)]
        tree = [polymul(tree[2 * r], tree[2 * r + 1]) for r in range(len(tree) // 2)]
        m <<= 1
    Z = [1]
    for x in range(n): Z = polymul(Z, [(-x) % p, 1])
    lu = n - 1
    u = [0] * max(lu, 1); u[0] = 1
    for i in range(1, lu):
        acc = sum(Z[n - t] * u[i - t] for t in range(1, i + 1)) % p
        u[i] = (-acc) % p
    Vblk = [x * invT % p for x in blocks_fwd(u[:lu], (lu + h - 1) // h, h, lgT)]
    return dict(TS=TS, h=h, lgT=lgT, Sb=Sb, nx=nx, ifact=ifact, Gblk=Gblk, Phat=Phat, Pnat=Pnat, Ptop=Ptop, Vblk=Vblk, invT=invT)

def interp_big(y, n, T):
    TS, h, lgT, Sb, nx = T['TS'], T['h'], T['lgT'], T['Sb'], T['nx']
    A = [y[i] * T['ifact'][i] % p if i < n else 0 for i in range(Sb)]
    X = blocks_fwd(A, nx, h, lgT)
    for k in range(nx):                       # Newton coefficients: low n of u * g
        Tk = block_out(X, nx, T['Gblk'], nx, k, TS, lgT)
        for i in range(TS):
            pos = k * h + i
            if pos >= Sb: continue
            if pos >= n: A[pos] = 0
            elif i >= h or k == 0: A[pos] = Tk[i]
            else: A[pos] = (A[pos] + Tk[i]) % p
    npad = (n + B - 1) // B * B
    for blk in range(npad // B):
        c = A[blk * B:(blk + 1) * B]
        for k in range(B - 2, -1, -1):
            pt = blk * B + k
            for j in range(k, B - 1): c[j] = (c[j] - c[j + 1] * pt) % p
        A[blk * B:(blk + 1) * B] = c
    m = B; lvl = 0
    Bf = [0] * Sb
    while m < n and 2 * m <= TS:
        lg = (2 * m).bit_length() - 1; two_m = 2 * m
        nb_active = (n - m + two_m - 1) // two_m; last = nb_active - 1
        h_last = min(m, n - (last * two_m + m)); shortp = h_last <= HMAX; nbN = nb_active - (1 if shortp else 0)
        for b in range(nbN):
            for i in range(m):
                v = A[b * two_m + m + i]; Bf[b * two_m + i] = v; Bf[b * two_m + m + i] = v
        hs = [A[last * two_m + m + i] for i in range(h_last)] if shortp else []
        if nbN:
            ntt_fwd(Bf, nbN, lg, 1)
            for idx in range(nbN * two_m): Bf[idx] = Bf[idx] * T['Phat'][lvl][idx] % p
            ntt_inv(Bf, nbN, lg)
            for idx in range(nbN * two_m):
                x = Bf[idx]
                if (idx & (two_m - 1)) < m: x = (x + A[idx]) % p
                A[idx] = x
        if shortp:
            Pn = T['Pnat'][lvl]
            for j in range(two_m):
                acc = 0; i = j - m if j > m else 0
                while i < h_last and i <= j: acc += hs[i] * Pn[j - i]; i += 1
                x = acc % p
                if j < m: x = (x + A[last * two_m + j]) % p
                A[last * two_m + j] = x
        m <<= 1; lvl += 1
    if m < n:                                  # the one level above TS: F = F_lo + x^m F_hi + p * F_hi, m = TS = 2h
        assert m == TS
        nf = (n - m + h - 1) // h
        X = blocks_fwd(A[m:], nf, h, lgT)
        for k in range(nf + 1):
            Tk = block_out(X, nf, T['Ptop'], 2, k, TS, lgT)
            for i in range(TS): A[k * h + i] = (A[k * h + i] + Tk[i]) % p
    return A[:n]

def quotient_big(a, b, n, T):
    TS, h, lgT, Sb, nx = T['TS'], T['h'], T['lgT'], T['Sb'], T['nx']
    X = blocks_fwd(a, nx, h, lgT); Y = blocks_fwd(b, nx, h, lgT)
    lu = n - 1
    U = [0] * Sb
    for k in range(2 * nx - 1):
        if k * h + TS <= n: continue
        Tk = block_out(X, nx, Y, nx, k, TS, lgT, T['invT'])
        for i in range(TS):
            pos = k * h + i
            if n <= pos <= 2 * n - 2: U[2 * n - 2 - pos] = (U[2 * n - 2 - pos] + Tk[i]) % p
    nu = (lu + h - 1) // h
    X = blocks_fwd(U[:lu], nu, h, lgT)
    R = [0] * Sb
    for k in range(nu):
        Tk = block_out(X, nu, T['Vblk'], nu, k, TS, lgT)
        for i in range(TS):
            pos = k * h + i
            if pos < lu: R[pos] = (R[pos] + Tk[i]) % p
    return [R[lu - 1 - i] for i in range(lu)]


@pytest.mark.parametrize("n,TS", [(48, 64), (50, 64), (64, 64), (65, 64), (80, 64), (81, 64), (97, 64), (127, 64), (128, 64),
                                  (150, 128), (256, 128), (300, 256), (600, 512), (1031, 1024), (1024, 512)])
def test_blocked_model_matches_oracle(n, TS):
    random.seed(n * 7 + TS)
    T = tables_big(n, TS)
    y = [random.randrange(p) for _ in range(n)]
    got = interp_big(y, n, T)
    want = O.interpolate(np.array(y, dtype=np.uint64).reshape(n, 1), 1, 1, [p])[:, 0]
    assert got == [int(x) for x in want]
    a = [random.randrange(p) for _ in range(n)]
    b = [random.randrange(p) for _ in range(n)]
    H = quotient_big(a, b, n, T)
    col = lambda v: np.array(v, dtype=np.uint64).reshape(n, 1)
    Hw, hl = O.witness_H(col(a), col(b), col([0] * n), 1, 1, [p])
    assert hl == n - 1 and H == [int(x) for x in Hw[:, 0]]
