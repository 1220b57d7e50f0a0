"""The C-ABI library loads (no GPU needed) and exports every symbol include/sa_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT
import __graft_entry__ as G


def test_library_exports_every_declared_symbol():
    G.build_cuda()
    G._paths()
    import sa_engine
    lib = sa_engine.load_library()
    header = open(os.path.join(ROOT, "include", "sa_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    bound = {name for name, _, _ in sa_engine.SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sa_b200.h but not exported"
        assert name in bound, f"{name} declared in sa_b200.h but not bound in sa_engine.SYMBOLS"
    assert lib.sa_version().startswith(b"sa_b200")


def test_host_alloc_without_a_gpu_fails_loudly():
    """sa_host_alloc needs the CUDA runtime for page-locked memory: on a box without a GPU it returns
    NULL and says why (there is no quiet fallback to malloc); with a GPU the buffer is usable"""
    G.build_cuda()
    G._paths()
    import sa_engine
    lib = sa_engine.load_library()
    p = lib.sa_host_alloc(1 << 20)
    if p is None:
        assert b"sa_host_alloc" in lib.sa_last_error()
    else:
        ctypes.memset(p, 0x5A, 1 << 20)
        assert lib.sa_host_free(p) == 0
    assert lib.sa_host_free(None) == 0


def test_marshal_roundtrip_and_pickle_identity():
    import pickle
    import random
    G.build_marshal()
    G._paths()
    import sa_marshal
    from hostmirror_loader import load_host_types
    T = load_host_types()
    rng = random.Random(2)
    xs = [T.fe(rng.randrange(T.field.p)) for _ in range(1000)] + [T.fe(0), T.fe(T.field.p - 1)]
    buf = sa_marshal.pack(xs)
    assert isinstance(buf, bytearray) and len(buf) == 16 * len(xs)
    import gc
    assert gc.isenabled()
    assert sa_marshal.unpack_ints(buf) == [x.value for x in xs]
    assert gc.isenabled(), "unpack_ints must leave the cyclic collector as it found it"
    sa_marshal.unpack(buf, xs[0].field, type(xs[0]))
    assert gc.isenabled(), "unpack must leave the cyclic collector as it found it"
    gc.disable()
    try:  # and a collector the caller had switched off stays off
        sa_marshal.unpack_ints(buf)
        assert not gc.isenabled()
    finally:
        gc.enable()
    ys = sa_marshal.unpack(buf, T.field, T.FieldElement)
    assert all(type(y) is T.FieldElement and y.field is T.field for y in ys)
    assert pickle.dumps(ys) == pickle.dumps(xs)
    assert sa_marshal.pack([x.value for x in xs]) == buf
