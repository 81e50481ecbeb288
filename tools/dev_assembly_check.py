"""Scratch GPU check: assembly parity vs the oracle (superseded by tests/test_gpu_assembly.py)."""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import isr_oracle as O
from isrsolver_b200 import _lib as L
lib = C.CDLL(L.LIB_PATH)
for name in ["isr_ctx_create","isr_problem_create","isr_assemble","isr_assemble_stats","isr_last_error","isr_kernel_kf"]:
    fn = getattr(lib, name); fn.restype, fn.argtypes = L.SIGNATURES[name]
ctx = C.c_void_p()
rc = lib.isr_ctx_create(0, C.byref(ctx)); print("ctx", rc, lib.isr_last_error())
def gpu_matrix(ecm, thr, settings, eff=None, ecm_err=None):
    n = len(ecm)
    st = np.ascontiguousarray([[int(a),b,c] for a,b,c in settings], dtype=np.int32)
    p = C.c_void_p()
    ed = None
    if eff is not None:
        ed = L.EpsDesc(); ed.kind = {"one":0,"row_table":1,"table_2d":2,"e_only":3}[eff.kind]
        ed.nx, ed.xlo, ed.xhi = eff.nx, eff.xlo, eff.xhi; ed.ne, ed.elo, ed.ehi = eff.ne, eff.elo, eff.ehi
        ed.values = L.dptr(eff.values)
        if eff.xedges is not None: ed.xedges = L.dptr(np.ascontiguousarray(eff.xedges))
        if eff.eedges is not None: ed.eedges = L.dptr(np.ascontiguousarray(eff.eedges))
    rc = lib.isr_problem_create(ctx, n, L.dptr(np.ascontiguousarray(ecm)), thr, len(st), L.iptr(st), C.byref(ed) if ed else None, C.byref(p))
    assert rc == 0, lib.isr_last_error()
    G = np.empty((n, n))
    ee = None if ecm_err is None else L.dptr(np.ascontiguousarray(ecm_err))
    t = time.time()
    rc = lib.isr_assemble(p, ee, L.dptr(G)); assert rc == 0, lib.isr_last_error()
    t1 = time.time() - t
    t = time.time()
    for _ in range(10): lib.isr_assemble(p, ee, L.dptr(G))
    t2 = (time.time() - t) / 10
    np_, nn = C.c_int64(), C.c_int64()
    lib.isr_assemble_stats(p, C.byref(np_), C.byref(nn))
    return G.T.copy(), t1, t2, np_.value
def cmp(name, G, Gref):
    scale = np.abs(Gref).max(axis=1, keepdims=True)
    rel = np.abs(G - Gref) / np.maximum(np.abs(Gref), 1e-300)
    relrow = np.abs(G - Gref) / scale
    big = np.abs(Gref) > 1e-6 * scale
    print(f"{name}: max rel (entries >1e-6 row max) {rel[big].max():.3e}; max err/rowmax {relrow.max():.3e}", flush=True)
# kernel spot values
xs = np.array([1e-6, 0.01, 0.5, 1e-3, 0.0013, 0.002]); ss = np.full_like(xs, 2.25); out = np.empty_like(xs)
lib.isr_kernel_kf(ctx, len(xs), L.dptr(xs), L.dptr(ss), L.dptr(out))
print("kf", out, [O.kernel_kf(x, 2.25) for x in xs])
for n in (20, 50):
    ecm = np.linspace(1.18, 2.0, n)
    for nm, st in (("linear", [(False,0,n-1)]), ("cspline", [(True,0,n-1)]), ("mixed", [(False,0,0),(True,1,n-1)]), ("mixed3", [(True,0,n//3),(False,n//3+1,n//2),(True,n//2+1,n-1)])):
        t=time.time(); Gref = O.Interp(ecm, 0.827, st).eval_eq_matrix(); to=time.time()-t
        G, t1, t2, npan = gpu_matrix(ecm, 0.827, st)
        print(f"N={n} {nm}: oracle {to:.2f}s gpu first {t1*1e3:.2f}ms steady {t2*1e3:.3f}ms panels {npan}")
        cmp(f"  N={n} {nm}", G, Gref)
n = 30; ecm = np.linspace(1.18, 2.0, n)
eff = O.row_table_efficiency(ecm, 512)
for nm, st in (("linear", [(False,0,n-1)]), ("mixed", [(False,0,0),(True,1,n-1)])):
    Gref = O.Interp(ecm, 0.827, st).eval_eq_matrix(eff=eff)
    G, t1, t2, npan = gpu_matrix(ecm, 0.827, st, eff=eff)
    print(f"eps512 N={n} {nm}: steady {t2*1e3:.3f}ms panels {npan}"); cmp("  eps512 "+nm, G, Gref)
# spread
ecm_err = np.full(n, 0.002)
for nm, st in (("linear", [(False,0,n-1)]), ("mixed", [(False,0,0),(True,1,n-1)])):
    Gref = O.Interp(ecm, 0.827, st).eval_eq_matrix(ecm_err=ecm_err)
    G, t1, t2, npan = gpu_matrix(ecm, 0.827, st, ecm_err=ecm_err)
    print(f"spread N={n} {nm}: steady {t2*1e3:.3f}ms"); cmp("  spread "+nm, G, Gref)
for n in (100, 200, 500):
    ecm = np.linspace(1.18, 2.0, n)
    for nm, st in (("linear", [(False,0,n-1)]), ("mixed", [(False,0,0),(True,1,n-1)])):
        G, t1, t2, npan = gpu_matrix(ecm, 0.827, st)
        print(f"N={n} {nm}: gpu first {t1*1e3:.2f}ms steady {t2*1e3:.3f}ms panels {npan} G00 {G[0,0]:.12f} G[n-1,n-1] {G[-1,-1]:.12f}")
    eff = O.row_table_efficiency(ecm, 512)
    G, t1, t2, npan = gpu_matrix(ecm, 0.827, [(False,0,n-1)], eff=eff)
    print(f"N={n} linear eps512: steady {t2*1e3:.3f}ms panels {npan}")
