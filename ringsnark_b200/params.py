"""Named parameter sets of the reference's configs, as plain integers so that bench.py and
the GPU tests need neither SEAL nor /root/reference at run time.  Values are what SEAL 4.1.1 produces
(CoeffModulus::Create / BFVDefault, depends/SEAL/native/src/seal/util/globals.cpp:23-71, first level = all but the
special prime); they are re-checked against the compiled reference in tests/test_params.py when oracle/_ref exists."""

# BFVDefault(8192) minus the special prime
Q_8192 = [8796092858369, 8796092792833, 17592186028033, 17592185438209]
# BFVDefault(16384) minus the special prime: 3 x 48-bit + 5 x 49-bit
Q_16384 = [281474976546817, 281474976317441, 281474975662081, 562949952798721, 562949952700417, 562949952274433,
           562949951979521, 562949951881217]

# first 8 primes of BFVDefault(32768) (55 bit): SURVEY.md 8(d) config C5 uses 8 data-level limbs at N_E = 2^15
Q_32768_8 = [0x7fffffffe90001, 0x7fffffffbf0001, 0x7fffffffbd0001, 0x7fffffffba0001, 0x7fffffffaa0001, 0x7fffffffa50001,
             0x7fffffff9f0001, 0x7fffffff7e0001]

CONFIGS = {
    # examples/example_SEAL.cpp: N_R = 4096, default_double_batching_modulus(4096, 8192) first level
    "c1": dict(N_R=4096, q=[68718428161, 68719230977], N_E=8192, Q=Q_8192, n=2, io=5, aux=1),
    # benchmarks/bench_mul_SEAL.cpp restated (N_R = 8192, N_E = 16384)
    "c3p": dict(N_R=8192, q=[8796092792833, 8796092858369, 17592183324673, 17592183390209], N_E=16384, Q=Q_16384,
                n=4, io=7, aux=1),
    # benchmarks/bench_logistic_regression_inference.cpp:20-27,72-126 (shape): one 54-bit ring prime
    "c4": dict(N_R=2048, q=[18014398508400641], N_E=16384, Q=Q_16384, n=1031, io=517, aux=1538),
    "c4m": dict(N_R=2048, q=[18014398508400641], N_E=16384, Q=Q_16384, n=129, io=65, aux=192),
    # SURVEY.md 8(d) C5 parameters (N_R = N_E = 2^15, one 54-bit ring prime = 1 mod 2^16, 8 x 55-bit limbs) with a circuit
    # small enough for the O(n^2) witness map; the n = 2^16 circuit itself needs the fast interpolation that is not built yet
    "c5s": dict(N_R=32768, q=[18014398506729473], N_E=32768, Q=Q_32768_8, n=257, io=129, aux=384),
    # C5 parameters with the largest circuit whose proving key (57 GiB), witness vectors and NTT-domain plaintexts fit ONE B200
    # (n = 2^16 needs the witness map beyond n = 16400 and a key of 1.3 TiB: DESIGN.md section 8)
    "c5m": dict(N_R=32768, q=[18014398506729473], N_E=32768, Q=Q_32768_8, n=4096, io=2049, aux=6144),
    "c4s": dict(N_R=2048, q=[18014398508400641], N_E=16384, Q=Q_16384, n=33, io=17, aux=48),
}


def _merge(x, y):
    """linear_combination::operator+ (relations/variable.tcc:268-300): sorted merge, equal indices add."""
    out, i, j = [], 0, 0
    while i < len(x) and j < len(y):
        if x[i][0] < y[j][0]:
            out.append(x[i]); i += 1
        elif x[i][0] > y[j][0]:
            out.append(y[j]); j += 1
        else:
            out.append((x[i][0], x[i][1] + y[j][1])); i += 1; j += 1
    return out + x[i:] + y[j:]


def synthetic_r1cs(n, io, aux, seed=1, use_const=False):
    """Satisfiable-shape R1CS wiring of oracle/cases.hpp::build_circuit (same generator, same order):
    (x_a [+ x_b] [+ 3]) * (x_c [+ 2 x_d]) = x_out.  Returns CSR (row_ptr[3n+1], col, coeff)."""
    M = (1 << 64) - 1
    s = (seed * 0x9E3779B97F4A7C15 + 0x1234567) & M

    def nxt():
        nonlocal s
        s ^= (s << 13) & M
        s ^= s >> 7
        s ^= (s << 17) & M
        return s

    nfree = io + aux - n
    rows = {0: [], 1: [], 2: []}
    for i in range(n):
        avail = nfree + i
        a, b, c, d = (nxt() % avail for _ in range(4))
        two_a = nxt() & 1
        two_b = (nxt() & 3) == 0
        konst = use_const and (nxt() % 3 == 0)
        ra = [(a + 1, 1)]
        if two_a:
            ra = _merge(ra, [(b + 1, 1)])
        if konst:
            ra = _merge(ra, [(0, 3)])
        rb = [(c + 1, 1)]
        if two_b:
            rb = _merge(rb, [(d + 1, 2)])
        rows[0].append(ra)
        rows[1].append(rb)
        rows[2].append([(nfree + i + 1, 1)])
    row_ptr, col, coeff = [0], [], []
    for m in range(3):
        for r in rows[m]:
            for (v, k) in r:
                col.append(v)
                coeff.append(k)
            row_ptr.append(len(col))
    return row_ptr, col, coeff
