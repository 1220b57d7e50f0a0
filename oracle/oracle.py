"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's NTT + FRI hot path:
a ctypes face over ``stark_oracle.c`` for bulk work plus pure-Python loops for
small cases.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as
the checker / the timed CPU arm.  The product (``stark-anatomy_b200/``) never
does.

Parity pin: ``tests/test_oracle.py`` checks every function here against
``tests/golden/*.json`` (made by importing the unmodified Python reference, see
``tests/golden/make_golden.py``) and against BASELINE.md section 3's 2^20 digests.

Citations: /root/reference/code/<file>:<lines>.
Bulk element layout: numpy ``uint64[n, 2]`` = (lo, hi) limbs of the canonical
residue, 16 bytes per element, little endian.
"""
import ctypes
import hashlib
import os
import pickle
import subprocess

import numpy as np

P = 1 + 407 * (1 << 119)                       # algebra.py:96-98
GENERATOR = 85408008396924667383611388730472331217  # algebra.py:100-102
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile libstark_oracle.so next to this file (gcc, OpenMP)."""
    so = os.path.join(_HERE, "libstark_oracle.so")
    src = os.path.join(_HERE, "stark_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libstark_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.so_init.restype = None
        L.so_num_threads.restype = ci
        L.so_set_threads.argtypes = [ci]
        L.so_set_threads.restype = None
        for name in ("so_fe_mul", "so_fe_add", "so_fe_sub", "so_fe_pow"):
            getattr(L, name).argtypes = [vp, vp, vp]
            getattr(L, name).restype = None
        L.so_fe_inv.argtypes = [vp, vp]
        L.so_ntt.argtypes = [vp, vp, sz, vp, ci]
        L.so_ntt.restype = ci
        L.so_intt.argtypes = [vp, vp, sz, vp, ci]
        L.so_intt.restype = ci
        L.so_ntt_batch.argtypes = [vp, vp, sz, sz, vp]
        L.so_ntt_batch.restype = ci
        L.so_pointwise_mul.argtypes = [vp, vp, vp, sz]
        L.so_pointwise_div.argtypes = [vp, vp, vp, sz]
        L.so_pointwise_div.restype = ci
        L.so_scale.argtypes = [vp, vp, sz, vp]
        L.so_poly_eval.argtypes = [vp, vp, sz, vp, sz]
        L.so_fri_fold.argtypes = [vp, vp, sz, vp, vp, vp]
        L.so_zerofier.argtypes = [vp, vp, sz]
        L.so_interpolate.argtypes = [vp, vp, vp, sz]
        L.so_interpolate.restype = ci
        L.so_blake2b.argtypes = [vp, vp, sz]
        L.so_decimal.argtypes = [vp, vp]
        L.so_decimal.restype = sz
        L.so_merkle_tree.argtypes = [vp, vp, sz]
        L.so_merkle_tree.restype = ci
        L.so_merkle_root.argtypes = [vp, vp, sz]
        L.so_merkle_root.restype = ci
        L.so_merkle_open.argtypes = [vp, vp, sz, sz]
        L.so_merkle_open.restype = ci
        L.so_init()
        _LIB = L
    return _LIB


# ------------------------------------------------------------ conversions --
def to_np(values):
    """list[int] -> uint64[n, 2]"""
    n = len(values)
    buf = b"".join(int(v).to_bytes(16, "little") for v in values)
    return np.frombuffer(buf, dtype="<u8").reshape(n, 2).copy()


def from_np(arr):
    """uint64[n, 2] -> list[int]"""
    raw = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    return [int.from_bytes(raw[i:i + 16], "little") for i in range(0, len(raw), 16)]


def _fe(v):
    return np.array([v & 0xFFFFFFFFFFFFFFFF, v >> 64], dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def vector_digest(arr):
    """BASELINE.md section 3 vector digest: blake2b over 16-byte LE values."""
    return hashlib.blake2b(np.ascontiguousarray(arr, dtype="<u8").tobytes()).hexdigest()


_ERR = {
    -1: "cannot compute ntt of non-power-of-two sequence",                                   # ntt.py:4
    -2: "primitive root must be nth root of unity, where n is len(values)",                  # ntt.py:10
    -3: "primitive root is not primitive nth root of unity, where n is len(values)",         # ntt.py:11
    -4: "divide by zero",                                                                    # algebra.py:92
}


# ------------------------------------------------------------------ field --
def primitive_nth_root(n):
    """algebra.py:104-114"""
    assert n <= 1 << 119 and (n & (n - 1)) == 0
    root, order = GENERATOR, 1 << 119
    while order != n:
        root = root * root % P
        order //= 2
    return root


def inverse(a):
    """algebra.py:87-89 (xgcd); inverse(0) == 0"""
    return pow(a, P - 2, P)


def sample(byte_array):
    """algebra.py:116-120 big-endian bytes -> int mod p"""
    return int.from_bytes(byte_array, "big") % P


# -------------------------------------------------------------------- ntt --
def ntt_np(root, arr, parallel=False):
    """ntt.py:3-18 on uint64[n,2]"""
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    out = np.empty_like(arr)
    rc = lib().so_ntt(_ptr(out), _ptr(arr), arr.shape[0], _ptr(_fe(root)), int(parallel))
    assert rc == 0, _ERR[rc]
    return out


def intt_np(root, arr, parallel=False):
    """ntt.py:20-30"""
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    out = np.empty_like(arr)
    rc = lib().so_intt(_ptr(out), _ptr(arr), arr.shape[0], _ptr(_fe(root)), int(parallel))
    assert rc == 0, _ERR[rc]
    return out


def ntt_batch_np(root, arr):
    """batch of independent ntt.py:3-18 transforms, uint64[B,n,2], OpenMP over B"""
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    out = np.empty_like(arr)
    rc = lib().so_ntt_batch(_ptr(out), _ptr(arr), arr.shape[1], arr.shape[0], _ptr(_fe(root)))
    assert rc == 0, _ERR[rc]
    return out


def ntt(root, values):
    return from_np(ntt_np(root, to_np(values))) if len(values) > 1 else values


def intt(root, values):
    return from_np(intt_np(root, to_np(values))) if len(values) > 1 else values


def py_ntt(root, values):
    """ntt.py:3-18 restated literally on Python ints (small cases only)."""
    n = len(values)
    assert n & (n - 1) == 0, _ERR[-1]
    if n <= 1:
        return values
    assert pow(root, n, P) == 1, _ERR[-2]
    assert pow(root, n // 2, P) != 1, _ERR[-3]
    half = n // 2
    odds = py_ntt(root * root % P, values[1::2])
    evens = py_ntt(root * root % P, values[::2])
    return [(evens[i % half] + pow(root, i, P) * odds[i % half]) % P for i in range(n)]


def pointwise_mul_np(a, b):
    out = np.empty_like(a)
    lib().so_pointwise_mul(_ptr(out), _ptr(a), _ptr(b), a.shape[0])
    return out


def pointwise_div_np(a, b):
    out = np.empty_like(a)
    rc = lib().so_pointwise_div(_ptr(out), _ptr(a), _ptr(b), a.shape[0])
    assert rc == 0, _ERR[rc]
    return out


def scale_np(arr, factor):
    """univariate.py:153-154"""
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    out = np.empty_like(arr)
    lib().so_scale(_ptr(out), _ptr(arr), arr.shape[0], _ptr(_fe(factor)))
    return out


def poly_eval_np(coeffs, points):
    """univariate.py:130-136 at every point"""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    points = np.ascontiguousarray(points, dtype=np.uint64)
    out = np.empty_like(points)
    lib().so_poly_eval(_ptr(out), _ptr(coeffs), coeffs.shape[0], _ptr(points), points.shape[0])
    return out


def zerofier_np(domain):
    """ntt.py:66-80: k+1 coefficients of prod (X - d)"""
    domain = np.ascontiguousarray(domain, dtype=np.uint64)
    out = np.empty((domain.shape[0] + 1, 2), dtype=np.uint64)
    lib().so_zerofier(_ptr(out), _ptr(domain), domain.shape[0])
    return out


def interpolate_np(domain, values):
    """ntt.py:102-130: k coefficients of the interpolant"""
    domain = np.ascontiguousarray(domain, dtype=np.uint64)
    values = np.ascontiguousarray(values, dtype=np.uint64)
    out = np.empty_like(domain)
    rc = lib().so_interpolate(_ptr(out), _ptr(domain), _ptr(values), domain.shape[0])
    assert rc == 0, _ERR[rc]
    return out


def degree(coeffs):
    """univariate.py:7-17"""
    d = -1
    for i, c in enumerate(coeffs):
        if c:
            d = i
    return d


def schoolbook_mul(l, r):
    """univariate.py:38-48 (untrimmed length len l + len r - 1)"""
    if not l or not r:
        return []
    buf = [0] * (len(l) + len(r) - 1)
    for i, a in enumerate(l):
        if a == 0:
            continue
        for j, b in enumerate(r):
            buf[i + j] = (buf[i + j] + a * b) % P
    return buf


def fast_multiply(lhs, rhs, root, order):
    """ntt.py:32-64 on coefficient lists of ints"""
    assert pow(root, order, P) == 1, "supplied root does not have supplied order"
    assert pow(root, order // 2, P) != 1, "supplied root is not primitive root of supplied order"
    dl, dr = degree(lhs), degree(rhs)
    if dl < 0 or dr < 0:
        return []
    deg = dl + dr
    if deg < 8:
        return schoolbook_mul(lhs, rhs)
    while deg < order // 2:
        root, order = root * root % P, order // 2
    a = lhs[:dl + 1] + [0] * (order - dl - 1)
    b = rhs[:dr + 1] + [0] * (order - dr - 1)
    prod = pointwise_mul_np(ntt_np(root, to_np(a)), ntt_np(root, to_np(b)))
    return from_np(intt_np(root, prod))[:deg + 1]


def fast_coset_evaluate(coeffs, offset, generator, order):
    """ntt.py:132-135"""
    scaled = scale_np(to_np(coeffs), offset) if coeffs else np.zeros((0, 2), np.uint64)
    padded = np.zeros((order, 2), dtype=np.uint64)
    padded[:len(coeffs)] = scaled
    return from_np(ntt_np(generator, padded))


def fast_coset_divide(lhs, rhs, offset, root, order):
    """ntt.py:137-176 (clean division only); degree < 8 long-division fallback is
    restated as exact division through the same coset route (identical values)."""
    assert pow(root, order, P) == 1, "supplied root does not have supplied order"
    assert pow(root, order // 2, P) != 1, "supplied root is not primitive root of supplied order"
    dl, dr = degree(lhs), degree(rhs)
    assert dr >= 0, "cannot divide by zero polynomial"
    if dl < 0:
        return []
    assert dr <= dl, "cannot divide by polynomial of larger degree"
    deg = max(dl, dr)
    while deg < order // 2:
        root, order = root * root % P, order // 2
    a = np.zeros((order, 2), np.uint64)
    b = np.zeros((order, 2), np.uint64)
    a[:dl + 1] = scale_np(to_np(lhs[:dl + 1]), offset)
    b[:dr + 1] = scale_np(to_np(rhs[:dr + 1]), offset)
    q = intt_np(root, pointwise_div_np(ntt_np(root, a), ntt_np(root, b)))
    return from_np(scale_np(q[:dl - dr + 1], inverse(offset)))


# ------------------------------------------------------------ merkle / fri --
def blake2b(msg):
    out = np.empty(64, dtype=np.uint8)
    m = np.frombuffer(bytes(msg), dtype=np.uint8) if len(msg) else np.zeros(1, np.uint8)
    lib().so_blake2b(_ptr(out), _ptr(m), len(msg))
    return out.tobytes()


def decimal(v):
    """algebra.py:53-57"""
    buf = np.empty(40, dtype=np.uint8)
    n = lib().so_decimal(_ptr(buf), _ptr(_fe(v)))
    return buf[:n].tobytes()


def merkle_tree_np(arr):
    """merkle.py:6-14; heap layout uint8[2n, 64], node 1 = root"""
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    n = arr.shape[0]
    tree = np.empty((2 * n, 64), dtype=np.uint8)
    rc = lib().so_merkle_tree(_ptr(tree), _ptr(arr), n)
    assert rc == 0, "length must be power of two"
    return tree


def merkle_root_np(arr):
    return merkle_tree_np(arr)[1].tobytes()


def merkle_open(tree, index):
    """merkle.py:16-27: list of sibling digests, bottom-up"""
    n = tree.shape[0] // 2
    assert 0 <= index < n, "cannot open invalid index"
    path = np.empty((max(n.bit_length() - 1, 1), 64), dtype=np.uint8)
    k = lib().so_merkle_open(_ptr(path), _ptr(tree), n, index)
    return [path[i].tobytes() for i in range(k)]


def fri_fold_np(arr, alpha, offset, omega):
    """fri.py:85"""
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    out = np.empty((arr.shape[0] // 2, 2), dtype=np.uint64)
    lib().so_fri_fold(_ptr(out), _ptr(arr), arr.shape[0], _ptr(_fe(alpha)), _ptr(_fe(offset)),
                      _ptr(_fe(omega)))
    return out


def fri_num_rounds(n, expansion_factor, num_colinearity_tests):
    """fri.py:22-28"""
    rounds = 0
    while n > expansion_factor and 4 * num_colinearity_tests < n:
        n //= 2
        rounds += 1
    return rounds


def fiat_shamir(objects, num_bytes=32):
    """ip.py:18-22"""
    return hashlib.shake_256(pickle.dumps(objects)).digest(num_bytes)


def fri_commit_np(codeword, offset, omega, expansion_factor, num_colinearity_tests,
                  prior_objects=()):
    """fri.py:56-96 with the transcript holding ``prior_objects`` then the roots.
    Returns (roots, alphas, layers) with layers as uint64[N_r, 2] arrays."""
    objects = list(prior_objects)
    rounds = fri_num_rounds(codeword.shape[0], expansion_factor, num_colinearity_tests)
    roots, alphas, layers = [], [], []
    cw = np.ascontiguousarray(codeword, dtype=np.uint64)
    for r in range(rounds):
        n = cw.shape[0]
        assert pow(omega, n - 1, P) == inverse(omega), "error in commit: omega does not have the right order!"
        root = merkle_root_np(cw)
        roots.append(root)
        objects.append(root)
        layers.append(cw)
        if r == rounds - 1:
            break
        alpha = sample(fiat_shamir(objects))
        alphas.append(alpha)
        cw = fri_fold_np(cw, alpha, offset, omega)
        omega, offset = omega * omega % P, offset * offset % P
    return roots, alphas, layers


def sample_indices(seed, size, reduced_size, number):
    """fri.py:30-51 (bytes(counter) is `counter` zero bytes)"""
    assert number <= reduced_size
    indices, reduced = [], []
    counter = 0
    while len(indices) < number:
        index = int.from_bytes(hashlib.blake2b(seed + bytes(counter)).digest(), "big") % size
        counter += 1
        if index % reduced_size not in reduced:
            indices.append(index)
            reduced.append(index % reduced_size)
    return indices
