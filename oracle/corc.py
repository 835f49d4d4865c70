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


def air_quotient(main_lde, aux_lde, log2_trace, offset, challenges, weights, mont_io=False):
    """all_quotients_combined on natural-order column-major LDE tables -> [N,3] (canonical in/out unless mont_io)"""
    m = np.ascontiguousarray(main_lde if mont_io else to_mont(main_lde)).reshape(main_lde.shape)
    a = np.ascontiguousarray(aux_lde if mont_io else to_mont(aux_lde)).reshape(aux_lde.shape)
    N = m.shape[1]
    ch, w = _xarr(challenges), _xarr(weights)
    assert ch.shape[0] == 63 and w.shape[0] == 604 and a.shape[1] == N
    out = np.empty((N, 3), dtype=np.uint64)
    lib().orc_air_quotient(_p(m), ctypes.c_size_t(m.shape[0]), _p(a), ctypes.c_size_t(a.shape[0]), ctypes.c_size_t(N),
                           ctypes.c_uint(log2_trace), ctypes.c_uint64(mont1(offset)), _p(ch), _p(w), _p(out))
    return out if mont_io else from_mont(out).reshape(N, 3)


def air_eval_category(cat, mc, ac, mn, an, challenges):
    """cat in 0..3; rows as canonical arrays (main 379 B-field, aux 90 X-field) -> [num_constraints, 3]"""
    l = lib()
    counts = (ctypes.c_int * 4).in_dll(l, "ORC_AIR_NUM_CONSTRAINTS")
    out = np.empty((counts[cat], 3), dtype=np.uint64)
    args = [to_mont(np.array(v, dtype=np.uint64).reshape(-1)) for v in (mc, ac, mn, an)]
    ch = _xarr(challenges)
    l.orc_air_eval_category(ctypes.c_int(cat), _p(args[0]), _p(args[1]), _p(args[2]), _p(args[3]), _p(ch), _p(out))
    return from_mont(out).reshape(-1, 3)


def fri_fold(cw, domain_offset, challenge, mont_io=False):
    c = np.ascontiguousarray(cw if mont_io else to_mont(np.array(cw, dtype=np.uint64).reshape(-1, 3))).reshape(-1, 3)
    n = c.shape[0]
    ch = _xarr([challenge])
    out = np.empty((n // 2, 3), dtype=np.uint64)
    lib().orc_fri_fold(_p(c), ctypes.c_size_t(n), ctypes.c_uint64(mont1(domain_offset)), _p(ch), _p(out))
    return out if mont_io else from_mont(out).reshape(-1, 3)


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


# ---- array stages of a complete prove (c/xpoly.c) -------------------------------------------------------------------
# Arrays stay in MONTGOMERY form between these calls (uint64 numpy, X-field arrays [k, 3]); `x3` converts one canonical
# X-field element.  Used by oracle/stark.py and oracle/stir.py when fast=True; the pure-Python code stays as cross-check.
def x3(x):
    return to_mont(np.array(x, dtype=np.uint64).reshape(3))

def x3_out(a):
    return tuple(int(v) for v in from_mont(np.ascontiguousarray(a).reshape(3)))

def xpoly_eval_m(c, x):
    c = np.ascontiguousarray(c).reshape(-1, 3)
    out = np.zeros(3, dtype=np.uint64)
    lib().orc_xpoly_eval(_p(c), ctypes.c_size_t(c.shape[0]), _p(x3(x)), _p(out))
    return x3_out(out)

def xpoly_fold_m(c, ff, r):
    c = np.ascontiguousarray(c).reshape(-1, 3)
    m = (c.shape[0] + ff - 1) // ff
    out = np.zeros((m, 3), dtype=np.uint64)
    if m:
        lib().orc_xpoly_fold(_p(c), ctypes.c_size_t(c.shape[0]), ctypes.c_size_t(ff), _p(x3(r)), _p(out))
    return out

def xpoly_mul_m(a, b):
    a = np.ascontiguousarray(a).reshape(-1, 3); b = np.ascontiguousarray(b).reshape(-1, 3)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return np.zeros((0, 3), dtype=np.uint64)
    out = np.zeros((a.shape[0] + b.shape[0] - 1, 3), dtype=np.uint64)
    lib().orc_xpoly_mul(_p(a), ctypes.c_size_t(a.shape[0]), _p(b), ctypes.c_size_t(b.shape[0]), _p(out))
    return out

def xpoly_trim_m(a):
    a = np.ascontiguousarray(a).reshape(-1, 3)
    nz = np.nonzero(a.any(axis=1))[0]
    return a[:nz[-1] + 1] if nz.size else a[:0]

def xpoly_div_m(num, den):
    num = xpoly_trim_m(num).copy(); den = xpoly_trim_m(den)
    if num.shape[0] < den.shape[0] or den.shape[0] == 0:
        return np.zeros((0, 3), dtype=np.uint64)
    q = np.zeros((num.shape[0] - den.shape[0] + 1, 3), dtype=np.uint64)
    lib().orc_xpoly_div(_p(num), ctypes.c_size_t(num.shape[0]), _p(den), ctypes.c_size_t(den.shape[0]), _p(q))
    return q

def xzerofier_m(points_m):
    p = np.ascontiguousarray(points_m).reshape(-1, 3)
    out = np.zeros((p.shape[0] + 1, 3), dtype=np.uint64)
    lib().orc_xzerofier(_p(p), ctypes.c_size_t(p.shape[0]), _p(out))
    return out

def xinterpolate_m(xs_m, ys_m):
    xs = np.ascontiguousarray(xs_m).reshape(-1, 3); ys = np.ascontiguousarray(ys_m).reshape(-1, 3)
    out = np.zeros((xs.shape[0], 3), dtype=np.uint64)
    lib().orc_xinterpolate(_p(xs), _p(ys), ctypes.c_size_t(xs.shape[0]), _p(out))
    return out

def xpoly_axpy_m(dst, src, w):
    """dst[:len(src)] += w * src (in place; dst at least as long as src)"""
    src = np.ascontiguousarray(src).reshape(-1, 3)
    assert dst.flags["C_CONTIGUOUS"] and dst.shape[0] >= src.shape[0]
    if src.shape[0]:
        lib().orc_xpoly_axpy(_p(dst), _p(src), ctypes.c_size_t(src.shape[0]), _p(x3(w)))

def xpoly_add_scaled_arg_m(dst, src, scale, arg):
    """dst[:len(src)] += scale * src(arg X) (B-field scale and arg, canonical ints)"""
    src = np.ascontiguousarray(src).reshape(-1, 3)
    assert dst.flags["C_CONTIGUOUS"] and dst.shape[0] >= src.shape[0]
    if src.shape[0]:
        lib().orc_xpoly_add_scaled_arg(_p(dst), _p(src), ctypes.c_size_t(src.shape[0]), ctypes.c_uint64(mont1(scale)), ctypes.c_uint64(mont1(arg)))

def interpolants_table_m(trace_m, rand_m):
    t = np.ascontiguousarray(trace_m)
    ncols, n = t.shape
    out = np.empty((ncols, 2 * n), dtype=np.uint64)
    h = rand_m.shape[1] if rand_m is not None else 0
    r = np.ascontiguousarray(rand_m) if rand_m is not None else None
    lib().orc_interpolants_table(_p(t), ctypes.c_uint(n.bit_length() - 1), ctypes.c_size_t(ncols), _p(r) if r is not None else None,
                                 ctypes.c_size_t(h), _p(out))
    return out

def bary_weights_m(log2n, alpha):
    dods = np.empty((1 << log2n, 3), dtype=np.uint64)
    di = np.zeros(3, dtype=np.uint64)
    lib().orc_bary_weights(ctypes.c_uint(log2n), _p(x3(alpha)), _p(dods), _p(di))
    return dods, di

def ood_row_m(cols_m, xf, dods, di, rand_m, alpha):
    """cols_m [xf * ncols, n] Montgomery, rand_m [xf * ncols, h] -> list of canonical X-field tuples"""
    c = np.ascontiguousarray(cols_m); r = np.ascontiguousarray(rand_m)
    ncols, n = c.shape[0] // xf, c.shape[1]
    out = np.zeros((ncols, 3), dtype=np.uint64)
    lib().orc_ood_row(_p(c), ctypes.c_size_t(ncols), ctypes.c_uint(n.bit_length() - 1), ctypes.c_int(xf), _p(dods), _p(di), _p(r),
                      ctypes.c_size_t(r.shape[1]), _p(x3(alpha)), _p(out))
    return [tuple(int(v) for v in row) for row in from_mont(out).reshape(-1, 3)]

def weighted_colsum_m(main_coef_m, aux_coef_m, weights):
    mc = np.ascontiguousarray(main_coef_m); ac = np.ascontiguousarray(aux_coef_m)
    w = _xarr(weights)
    ln = mc.shape[1]
    assert ac.shape[1] == ln and w.shape[0] == mc.shape[0] + ac.shape[0] // 3
    out = np.empty((ln, 3), dtype=np.uint64)
    lib().orc_weighted_colsum(_p(mc), ctypes.c_size_t(mc.shape[0]), _p(ac), ctypes.c_size_t(ac.shape[0] // 3), ctypes.c_size_t(ln), _p(w), _p(out))
    return out

def deep_combination_m(cw_ma, cw_p, cw_r, offset, points, values, weights):
    a, b, c = (np.ascontiguousarray(v).reshape(-1, 3) for v in (cw_ma, cw_p, cw_r))
    N = a.shape[0]
    out = np.empty((N, 3), dtype=np.uint64)
    lib().orc_deep_combination(_p(a), _p(b), _p(c), ctypes.c_uint(N.bit_length() - 1), ctypes.c_uint64(mont1(offset)), _p(_xarr(points)),
                               _p(_xarr(values)), _p(_xarr(weights)), _p(out))
    return out

def xcoset_evaluate_m(coeffs_m, offset, n):
    """X-field polynomial (Montgomery [k,3], k <= n) -> evaluations on offset * <w_n>, Montgomery [n,3]"""
    c = np.ascontiguousarray(coeffs_m).reshape(-1, 3)
    out = np.zeros((n, 3), dtype=np.uint64)
    if c.shape[0] == 0:
        return out
    assert c.shape[0] <= n
    log2n = n.bit_length() - 1
    om = ctypes.c_uint64(mont1(offset))
    for d in range(3):
        col = np.ascontiguousarray(c[:, d])
        o = np.empty(n, dtype=np.uint64)
        lib().orc_coset_evaluate(_p(col), ctypes.c_size_t(col.size), om, ctypes.c_uint(log2n), _p(o))
        out[:, d] = o
    return out

def xcoset_interpolate_m(values_m, offset):
    v = np.ascontiguousarray(values_m).reshape(-1, 3)
    n = v.shape[0]
    out = np.empty((n, 3), dtype=np.uint64)
    om = ctypes.c_uint64(mont1(offset))
    for d in range(3):
        col = np.ascontiguousarray(v[:, d])
        o = np.empty(n, dtype=np.uint64)
        lib().orc_coset_interpolate(_p(col), om, ctypes.c_uint(n.bit_length() - 1), _p(o))
        out[:, d] = o
    return out
