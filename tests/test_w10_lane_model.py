"""Exact integer model of the lane program in secp256k1-zkp_b200/csrc/ge_w10.cuh (one 26-bit limb of every field element per lane): every
intermediate is asserted to fit the register width the CUDA code gives it (u32 / u64, differences non-negative), on random inputs and on the
extreme magnitudes the doubling / addition formulas produce, and the results are compared with plain affine arithmetic mod p.  The CUDA
kernels themselves are compared with the reference on the GPU (tests/test_gpu_round2.py, unit ops 27-29); this file pins the arithmetic
design -- column split, the three folds through 2^260 = 2^36 + 0x3D10 and 2^256 = 2^32 + 0x3D1, carry rounds, a + 16p - b."""
import random

P = 2**256 - 2**32 - 977
M26 = (1 << 26) - 1
M22 = (1 << 22) - 1
PL = [0x3FFFC2F, 0x3FFFFBF] + [0x3FFFFFF] * 7 + [0x03FFFFF]


def val(l):
    return sum(v << (26 * i) for i, v in enumerate(l))


def u32(x):
    assert 0 <= x < 2**32, x
    return x


def u64(x):
    assert 0 <= x < 2**64, x
    return x


def mul(a, b):                                   # w10_mul
    lo = [0] * 10; hi = [0] * 10
    for k in range(10):
        for i in range(10):
            j = (k - i) % 10
            if i <= k:
                lo[k] += a[i] * b[j]
            else:
                hi[k] += a[i] * b[j]
        u64(lo[k]); u64(hi[k])
    assert hi[9] == 0
    hl = [h & M26 for h in hi]; hh = [u32(h >> 26) for h in hi]
    e = [u32(hl[k] + hh[(k + 9) % 10]) for k in range(10)]
    d = []
    for k in range(10):
        x = lo[k] + e[k] * 0x3D10 + (e[(k + 9) % 10] * 0x400 if k else 0)
        x += e[9] * (0xF44000 if k == 0 else 0x100000 if k == 1 else 0)
        d.append(u64(x))
    l = [x & M26 for x in d]; h = [u32(x >> 26) for x in d]
    r = [u64(l[k] + (h[(k + 9) % 10] if k else h[9] * 0x3D10) + (h[9] * 0x400 if k == 1 else 0)) for k in range(10)]
    l2 = [x & (M22 if k == 9 else M26) for k, x in enumerate(r)]
    c = [u32(x >> (22 if k == 9 else 26)) for k, x in enumerate(r)]
    return [u32(l2[k] + (c[(k + 9) % 10] if k else u32(c[9] * 0x3D1)) + (u32(c[9] * 0x40) if k == 1 else 0)) for k in range(10)]


def norm(v):                                     # w10_norm
    for x in v:
        u32(x)
    l = [x & (M22 if k == 9 else M26) for k, x in enumerate(v)]
    c = [x >> (22 if k == 9 else 26) for k, x in enumerate(v)]
    return [u32(l[k] + (c[(k + 9) % 10] if k else u32(c[9] * 0x3D1)) + (u32(c[9] * 0x40) if k == 1 else 0)) for k in range(10)]


def sub16(a, b, extra=None):                     # a + 16 p - b (- extra), limb by limb, never negative
    out = []
    for k in range(10):
        v = a[k] + 16 * PL[k]
        s = b[k] + (extra[k] if extra else 0)
        assert v >= s
        out.append(u32(v - s))
    return out


def from_int(x):
    return [(x >> (26 * i)) & M26 for i in range(10)]


def dbl(X, Y, Z):                                # w10_double_n, one round
    A = mul(X, X); B = mul(Y, Y); YZ = mul(Y, Z)
    XB = mul(X, B); C = mul(B, B); A2 = mul(A, A)
    X3 = norm(sub16([9 * v for v in A2], [8 * v for v in XB]))
    Nn = norm(sub16([12 * v for v in XB], [9 * v for v in A2]))
    M = mul([u32(3 * v) for v in A], Nn)
    return X3, norm(sub16(M, [8 * v for v in C])), [u32(2 * v) for v in YZ]


def add(Xa, Ya, Za, Xb, Yb, Zb):                 # w10_add, generic branch
    z22 = mul(Zb, Zb); z12 = mul(Za, Za); zab = mul(Za, Zb)
    u1 = mul(Xa, z22); u2 = mul(Xb, z12); t1 = mul(Zb, z22)
    H = norm(sub16(u2, u1))
    t2 = mul(Za, z12); S1 = mul(Ya, t1); HH = mul(H, H)
    S2 = mul(Yb, t2); HHH = mul(H, HH); V = mul(u1, HH)
    R = norm(sub16(S2, S1))
    RR = mul(R, R); S1H = mul(S1, HHH); Z3 = mul(zab, H)
    X3 = norm(sub16(RR, HHH, [2 * v for v in V]))
    D = norm(sub16(V, X3))
    return X3, norm(sub16(mul(R, D), S1H)), Z3, val(H) % P == 0, val(R) % P == 0


def aff(X, Y, Z):
    x, y, z = val(X) % P, val(Y) % P, val(Z) % P
    if z == 0:
        return None
    zi = pow(z, -1, P)
    return (x * zi * zi % P, y * zi * zi * zi % P)


def ec_dbl(p):
    x, y = p
    lam = 3 * x * x * pow(2 * y, -1, P) % P
    x3 = (lam * lam - 2 * x) % P
    return (x3, (lam * (x - x3) - y) % P)


def ec_add(p, q):
    if p[0] == q[0]:
        return None if (p[1] + q[1]) % P == 0 else ec_dbl(p)
    lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, P) % P
    x3 = (lam * lam - p[0] - q[0]) % P
    return (x3, (lam * (p[0] - x3) - p[1]) % P)


def lift(x):
    while True:
        y2 = (x**3 + 7) % P
        y = pow(y2, (P + 1) // 4, P)
        if y * y % P == y2:
            return (x, y)
        x += 1


def jac(pt, z):
    return [from_int(pt[0] * z * z % P), from_int(pt[1] * z**3 % P), from_int(z)]


def test_multiplication_bounds_and_value():
    rnd = random.Random(1)
    top = (1 << 26) + (1 << 21)
    for it in range(3000):
        if it % 4 == 0:                          # extreme magnitudes: a at 1, b at 4 (the largest product the formulas form)
            a = [rnd.choice([0, top, M26, 1]) for _ in range(10)]
            b = [rnd.choice([0, 4 * top - 1, M26]) for _ in range(10)]
        else:
            a = from_int(rnd.randrange(P)); b = from_int(rnd.randrange(P))
        s = mul(a, b)
        assert val(s) % P == val(a) * val(b) % P
        assert all(s[k] < top for k in range(9)) and s[9] < (1 << 22) + (1 << 7)


def test_doubling_chain_and_addition():
    rnd = random.Random(2)
    for it in range(40):
        a = lift(rnd.randrange(P)); b = lift(rnd.randrange(P))
        A = jac(a, rnd.randrange(1, P)); B = jac(b, rnd.randrange(1, P))
        ra, rb = a, b
        for _ in range(it % 20):                 # the invariant between operations: X, Y magnitude 1, Z magnitude <= 2
            A = list(dbl(*A)); ra = ec_dbl(ra)
            assert aff(*A) == ra
        if it % 3 == 0:
            B = list(dbl(*B)); rb = ec_dbl(rb)
        X3, Y3, Z3, hz, rz = add(*A, *B)
        assert not hz and aff(X3, Y3, Z3) == ec_add(ra, rb)
        X, Y, Z = dbl(X3, Y3, Z3)                # addition results feed doublings and additions again
        X, Y, Z, _, _ = add(X, Y, Z, *B)
        assert aff(X, Y, Z) == ec_add(ec_dbl(ec_add(ra, rb)), rb)
    # the degenerate branches are told apart by exact zero tests of H and R
    a = lift(12345)
    A = jac(a, 77)
    assert add(*A, *jac(a, 2**200 + 5))[3:] == (True, True)                              # same point, different z: doubling
    assert add(*A, *jac((a[0], P - a[1]), 2**200 + 5))[3:] == (True, False)              # opposite points: infinity
    X, Y, Z = dbl(from_int(5), from_int(7), from_int(0))                                 # infinity stays infinity (Z exactly zero)
    assert val(Z) == 0
