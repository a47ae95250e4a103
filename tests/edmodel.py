"""TEST INFRASTRUCTURE: tiny big-integer model of edwards25519 used only to CONSTRUCT adversarial
inputs (small-order / mixed-order / non-canonical points, honest signatures for them).  Expected
results never come from this file — they come from the compiled reference / the oracle."""
import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
SQRTM1 = pow(2, (P - 1) // 4, P)


def inv(x):
    return pow(x, P - 2, P)


def add(p1, p2):
    (x1, y1), (x2, y2) = p1, p2
    k = D * x1 * x2 * y1 * y2 % P
    return ((x1 * y2 + x2 * y1) * inv(1 + k) % P, (y1 * y2 + x1 * x2) * inv(1 - k) % P)


def mul(k, pt):
    r = (0, 1)
    while k > 0:
        if k & 1:
            r = add(r, pt)
        pt = add(pt, pt)
        k >>= 1
    return r


def neg(pt):
    return ((-pt[0]) % P, pt[1])


def recover_x(y, sign):
    """x with x^2 = (y^2-1)/(d y^2+1) and lsb == sign, or None if y is not on the curve."""
    u = (y * y - 1) % P
    v = (D * y * y + 1) % P
    x2 = u * inv(v) % P
    x = pow(x2, (P + 3) // 8, P)
    if (x * x - x2) % P != 0:
        x = x * SQRTM1 % P
    if (x * x - x2) % P != 0:
        return None
    if x & 1 != sign:
        x = (-x) % P
    return x


BY = 4 * inv(5) % P
B = (recover_x(BY, 0), BY)


def enc(pt, noncanonical=False, flip_sign=False):
    x, y = pt
    if noncanonical:
        assert y < 19
        y += P
    v = y | ((x & 1) << 255)
    if flip_sign:
        v ^= 1 << 255
    return v.to_bytes(32, "little")


def on_curve_y(ybytes):
    v = int.from_bytes(ybytes, "little")
    y = (v & ((1 << 255) - 1)) % P
    return recover_x(y, 0) is not None


def torsion8():
    """A point of exact order 8."""
    y = 2
    while True:
        x = recover_x(y, 0)
        if x is not None:
            t = mul(L, (x, y))
            if mul(4, t) != (0, 1):
                return t
        y += 1


def small_order_points():
    t = torsion8()
    return [mul(k, t) for k in range(8)]


def h_mod_l(*parts):
    return int.from_bytes(hashlib.sha512(b"".join(parts)).digest(), "little") % L


def clamp_scalar(sk):
    h = bytearray(hashlib.sha512(sk).digest())
    h[0] &= 0xF8
    h[31] &= 0x7F
    h[31] |= 0x40
    return int.from_bytes(h[:32], "little"), bytes(h[32:])


def sign_with(a, prefix, A_bytes, msg):
    """RFC 8032-style signature for secret scalar a and arbitrary public-key bytes A_bytes."""
    r = h_mod_l(prefix, msg)
    R = enc(mul(r, B))
    t = h_mod_l(R, A_bytes, msg)
    s = (r + t * a) % L
    return R + s.to_bytes(32, "little")


def check_comb_table(raw, w):
    """raw: (rows, entries, 96) uint8 — the fixed-base comb table; asserts entry [j][k] == (k + 1) * 2^(w j) * B as
    (y+x, y-x, 2dxy), canonical little-endian, for every entry."""
    rows, entries = raw.shape[0], raw.shape[1]
    base = B
    for j in range(rows):
        acc = base
        for k in range(entries):
            x, y = acc
            want = b"".join(v.to_bytes(32, "little") for v in ((y + x) % P, (y - x) % P, 2 * D * x * y % P))
            assert raw[j, k].tobytes() == want, (j, k)
            acc = add(acc, base)
        for _ in range(w):
            base = add(base, base)
