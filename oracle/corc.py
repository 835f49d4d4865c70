"""ctypes binding of the C oracle (oracle/c/tvm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Arrays cross this binding as numpy uint64 in CANONICAL form unless `mont=True`;
the C side computes in Montgomery form."""
import ctypes, os, subprocess
import numpy as np
from . import tip5 as _tip5

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libtvm_oracle.so")
_U64P = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    srcs = [os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c"))]   # .c, .h and generated .inc
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libtvm_oracle.so"] + (["-B"] if force else []))
    return _LIB


_lib = None
def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_mul.restype = ctypes.c_uint64
        _lib.orc_mul.argtypes = [ctypes.c_uint64] * 2
        _lib.orc_pow.restype = ctypes.c_uint64
        _lib.orc_pow.argtypes = [ctypes.c_uint64] * 2
        _lib.orc_inv.restype = ctypes.c_uint64
        _lib.orc_inv.argtypes = [ctypes.c_uint64]
        _lib.orc_root_of_unity.restype = ctypes.c_uint64
        rc = (ctypes.c_uint64 * 80)(*_tip5.ROUND_CONSTANTS)
        _lib.orc_tip5_set_round_constants(rc)
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_U64P)

def to_mont(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_to_mont(_p(a), ctypes.c_size_t(a.size))
    return a

def from_mont(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_from_mont(_p(a), ctypes.c_size_t(a.size))
    return a

def mont1(x):
    return int(to_mont(np.array([x % _tip5.P], dtype=np.uint64))[0])

def ntt(x, inverse=False):
    """canonical in/out, natural order"""
    a = to_mont(x)
    log2n = a.size.bit_length() - 1
    (lib().orc_intt if inverse else lib().orc_ntt)(_p(a), ctypes.c_uint(log2n))
    return from_mont(a)

def xntt(x, inverse=False):
    """x: [n,3] canonical"""
    a = to_mont(x)
    log2n = (a.size // 3).bit_length() - 1
    (lib().orc_xintt if inverse else lib().orc_xntt)(_p(a), ctypes.c_uint(log2n))
    return from_mont(a).reshape(-1, 3)

def coset_evaluate(coef, offset, log2n):
    c = to_mont(coef)
    out = np.empty(1 << log2n, dtype=np.uint64)
    lib().orc_coset_evaluate(_p(c), ctypes.c_size_t(c.size), ctypes.c_uint64(mont1(offset)), ctypes.c_uint(log2n), _p(out))
    return from_mont(out)

def coset_interpolate(vals, offset):
    v = to_mont(vals)
    out = np.empty_like(v)
    lib().orc_coset_interpolate(_p(v), ctypes.c_uint64(mont1(offset)), ctypes.c_uint(v.size.bit_length() - 1), _p(out))
    return from_mont(out)

def lde_table(trace_colmajor, randomizers, eval_offset, log2_eval, mont_io=False):
    """trace_colmajor [ncols, n]; randomizers [ncols, h] or None -> [ncols, 2^log2_eval]"""
    t = trace_colmajor if mont_io else to_mont(trace_colmajor)
    t = np.ascontiguousarray(t)
    ncols, n = t.shape
    r = None
    if randomizers is not None:
        r = np.ascontiguousarray(randomizers if mont_io else to_mont(randomizers)).reshape(ncols, -1)
    out = np.empty((ncols, 1 << log2_eval), dtype=np.uint64)
    lib().orc_lde_table(_p(t), ctypes.c_uint(n.bit_length() - 1), ctypes.c_size_t(ncols),
                        _p(r) if r is not None else None, ctypes.c_size_t(r.shape[1] if r is not None else 0),
                        ctypes.c_uint64(mont1(eval_offset)), ctypes.c_uint(log2_eval), _p(out))
    return out if mont_io else from_mont(out).reshape(out.shape)

def permutation(state):
    s = to_mont(np.array(state, dtype=np.uint64))
    lib().orc_tip5_permutation(_p(s))
    return [int(v) for v in from_mont(s)]

def hash_varlen(words):
    w = to_mont(np.array(words, dtype=np.uint64))
    d = np.empty(5, dtype=np.uint64)
    lib().orc_hash_varlen(_p(w), ctypes.c_size_t(w.size), _p(d))
    return [int(v) for v in from_mont(d)]

def hash_rows_colmajor(table, mont_io=False):
    """table [ncols, nrows] -> digests [nrows, 5]"""
    t = np.ascontiguousarray(table if mont_io else to_mont(table)).reshape(table.shape)
    ncols, nrows = t.shape
    d = np.empty((nrows, 5), dtype=np.uint64)
    lib().orc_hash_rows_colmajor(_p(t), ctypes.c_size_t(nrows), ctypes.c_size_t(ncols), ctypes.c_size_t(nrows), _p(d))
    return d if mont_io else from_mont(d).reshape(d.shape)

def merkle_build(leaves, mont_io=False):
    """leaves [n,5] -> nodes [2n,5] (nodes[1] = root)"""
    l = np.ascontiguousarray(leaves if mont_io else to_mont(leaves)).reshape(leaves.shape)
    n = l.shape[0]
    nodes = np.empty((2 * n, 5), dtype=np.uint64)
    lib().orc_merkle_build(_p(l), ctypes.c_size_t(n), _p(nodes))
    return nodes if mont_io else from_mont(nodes).reshape(nodes.shape)

def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def _xarr(xs):
    """list of X-field tuples / [k,3] array -> Montgomery uint64 [k,3]"""
    return to_mont(np.array(xs, dtype=np.uint64).reshape(-1, 3)).reshape(-1, 3)


def air_quotient(main_lde, aux_lde, log2_trace, offset, challenges, weights):
    """all_quotients_combined on natural-order column-major LDE tables -> [N,3] canonical"""
    m = np.ascontiguousarray(to_mont(main_lde)).reshape(main_lde.shape)
    a = np.ascontiguousarray(to_mont(aux_lde)).reshape(aux_lde.shape)
    N = m.shape[1]
    ch, w = _xarr(challenges), _xarr(weights)
    assert ch.shape[0] == 63 and w.shape[0] == 604 and a.shape[1] == N
    out = np.empty((N, 3), dtype=np.uint64)
    lib().orc_air_quotient(_p(m), ctypes.c_size_t(m.shape[0]), _p(a), ctypes.c_size_t(a.shape[0]), ctypes.c_size_t(N),
                           ctypes.c_uint(log2_trace), ctypes.c_uint64(mont1(offset)), _p(ch), _p(w), _p(out))
    return from_mont(out).reshape(N, 3)


def air_eval_category(cat, mc, ac, mn, an, challenges):
    """cat in 0..3; rows as canonical arrays (main 379 B-field, aux 90 X-field) -> [num_constraints, 3]"""
    l = lib()
    counts = (ctypes.c_int * 4).in_dll(l, "ORC_AIR_NUM_CONSTRAINTS")
    out = np.empty((counts[cat], 3), dtype=np.uint64)
    args = [to_mont(np.array(v, dtype=np.uint64).reshape(-1)) for v in (mc, ac, mn, an)]
    ch = _xarr(challenges)
    l.orc_air_eval_category(ctypes.c_int(cat), _p(args[0]), _p(args[1]), _p(args[2]), _p(args[3]), _p(ch), _p(out))
    return from_mont(out).reshape(-1, 3)


def fri_fold(cw, domain_offset, challenge):
    c = np.ascontiguousarray(to_mont(np.array(cw, dtype=np.uint64).reshape(-1, 3))).reshape(-1, 3)
    n = c.shape[0]
    ch = _xarr([challenge])
    out = np.empty((n // 2, 3), dtype=np.uint64)
    lib().orc_fri_fold(_p(c), ctypes.c_size_t(n), ctypes.c_uint64(mont1(domain_offset)), _p(ch), _p(out))
    return from_mont(out).reshape(-1, 3)


def aux_extend(main_table, challenges, randomizer_column=None):
    """MasterMainTable::extend (master_table.rs:1006-1075) through the rules generated from the AIR
    (triton-vm_b200/airgen/extend_gen.py -> c/aux_extend_gen.inc), run as a sequential loop.
    main_table: [379][n] canonical; challenges: 63 X-field triples; randomizer_column: [n][3] canonical (column 90).
    -> [91][n][3] canonical"""
    T = to_mont(np.ascontiguousarray(main_table, dtype=np.uint64))
    ncols, n = T.shape
    assert ncols == 379
    ch = to_mont(np.array(challenges, dtype=np.uint64).reshape(63, 3))
    aux = np.zeros((91 * 3, n), dtype=np.uint64)
    if randomizer_column is not None:
        aux[270:273] = to_mont(np.ascontiguousarray(np.asarray(randomizer_column, dtype=np.uint64).reshape(n, 3).T))
    lib().orc_aux_extend(_p(T), ctypes.c_size_t(n), _p(ch), _p(aux))
    return np.ascontiguousarray(from_mont(aux).reshape(91, 3, n).transpose(0, 2, 1))


def fill_derived_main(main_table):
    """DegreeLoweringTable::fill_derived_main_columns through the generated rules: [379][n] canonical -> copy with columns
    149..378 recomputed from columns 0..148"""
    T = to_mont(np.ascontiguousarray(main_table, dtype=np.uint64))
    assert T.shape[0] == 379
    lib().orc_fill_derived_main(_p(T), ctypes.c_size_t(T.shape[1]))
    return from_mont(T).reshape(main_table.shape)
