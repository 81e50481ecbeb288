"""Scratch GPU check: factor + toy batch parity and timing."""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import isr_oracle as O
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
def problem(ecm, thr, settings):
    st = np.ascontiguousarray([[int(a),b,c] for a,b,c in settings], dtype=np.int32); p = C.c_void_p()
    L.check(lib.isr_problem_create(ctx, len(ecm), L.dptr(np.ascontiguousarray(ecm)), thr, len(st), L.iptr(st), None, C.byref(p)))
    return p
for n, T in ((30, 1000), (50, 5000), (101, 3000), (200, 20000)):
    ecm = np.linspace(1.18, 2.0, n); st = [(False,0,0),(True,1,n-1)]
    p = problem(ecm, 0.827, st)
    G = np.empty((n,n)); L.check(lib.isr_assemble(p, None, L.dptr(G))); G = G.T.copy()
    bcs0 = O.bw_born(ecm); vcs = G @ bcs0; err = 0.05*vcs
    L.check(lib.isr_factor(p, L.dptr(err)))
    b = np.empty(n); cov = np.empty((n,n)); L.check(lib.isr_solve(p, L.dptr(vcs), L.dptr(b), L.dptr(cov)))
    bo, covo = O.sle_solve(G, vcs, err)
    print(f"N={n}: solve rel {np.abs(b-bo).max()/np.abs(bo).max():.2e} cov rel {np.abs(cov.T-covo).max()/np.abs(covo).max():.2e}")
    V = O.draw_toys(vcs, err, T)
    chi2 = np.empty(T); sb = np.empty(n); rs = np.empty(3*n); rh = np.zeros((n,1026), dtype=np.uint32); B = np.empty((T,n))
    L.check(lib.isr_toy_batch(p, T, L.dptr(V), L.dptr(bcs0), L.dptr(chi2), L.dptr(sb), L.dptr(rs), L.uptr(rh), L.dptr(B)))
    c2, Bo = O.chi2_test_fair(G, V, bcs0, err)
    means, errs, counts, stats = O.ratio_test(Bo, bcs0)
    print(f"   chi2 rel {np.abs(chi2/c2-1).max():.2e} B rel {np.abs(B-Bo).max()/np.abs(Bo).max():.2e} sumb rel {np.abs(sb-Bo.sum(0)).max()/np.abs(Bo.sum(0)).max():.2e}"
          f" stats rel {np.abs(rs.reshape(n,3)-stats).max():.2e} hist mismatches {(rh != counts.astype(np.uint32)).sum()}")
    if n <= 50:
        c2l, Bl = O.chi2_test_loop(G, V[:50], bcs0, err)
        print(f"   vs faithful loop (50 toys): chi2 rel {np.abs(chi2[:50]/c2l-1).max():.2e}")
    lo, hi = C.c_double(), C.c_double(); cnt = np.zeros(130, dtype=np.uint32)
    L.check(lib.isr_chi2_histogram(ctx, T, L.dptr(chi2), 128, C.byref(lo), C.byref(hi), L.uptr(cnt)))
    co, idx, lo2, hi2 = O.chi2_histogram(chi2)
    print("   hist equal", np.array_equal(cnt, co.astype(np.uint32)), lo.value == lo2, hi.value == hi2)
# timing, device-resident
stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx))
for n in (100, 200, 500):
    ecm = np.linspace(1.18, 2.0, n); st = [(False,0,0),(True,1,n-1)]
    p = problem(ecm, 0.827, st)
    G = np.empty((n,n)); L.check(lib.isr_assemble(p, None, L.dptr(G))); G = G.T.copy()
    bcs0 = O.bw_born(ecm); vcs = G @ bcs0; err = 0.05*vcs
    L.check(lib.isr_factor(p, L.dptr(err)))
    T = (1 << 20) if n <= 200 else (1 << 18)
    Vd = torch.from_numpy(vcs)[None,:].cuda() + torch.from_numpy(err)[None,:].cuda()*torch.randn(T, n, dtype=torch.float64, device='cuda')
    chi2d = torch.empty(T, dtype=torch.float64, device='cuda'); sbd = torch.zeros(n, dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    for cfgname in ({100: [None, "m128n112w16", "m128n112w16s2k40"], 200: [None, "m64n200w20", "m64n200w20s2k40"], 500: [None, "m64n256w16", "m64n256w16s2k40"]}[n]):
        if cfgname: os.environ["ISR_TOY_CFG"] = cfgname
        else: os.environ.pop("ISR_TOY_CFG", None)
        for modes in ("chi2+sum", "chi2"):
            sb = sbd.data_ptr() if modes == "chi2+sum" else None
            with torch.cuda.stream(stream):
                for _ in range(2): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), sb, None, None, None))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(3): L.check(lib.isr_toy_batch_dev(p, T, Vd.data_ptr(), L.dptr(bcs0), chi2d.data_ptr(), sb, None, None, None))
                e1.record(stream)
            e1.synchronize(); ms = e0.elapsed_time(e1)/3
            print(f"N={n} T={T} cfg={cfgname} {modes}: {ms:.3f} ms  {T/ms*1e3:.3e} toys/s  {4*n*n*T/ms*1e-9:.2f} TFLOP/s  mean chi2 {chi2d.mean().item():.3f}")
