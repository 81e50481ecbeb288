"""Critical-chain timing of the cluster Gauss-Jordan kernel (needs a build with -DISR_GJ_TRACE, see tools/README.md):
per pivot block, clock64 stamps of the owner warp (start / credits ok / eliminated / published) and of the warp that owns
the NEXT block (before wait / block visible / applied)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from isrsolver_b200 import _lib as L
lib = L.load()
ctx = C.c_void_p(); L.check(lib.isr_ctx_create(0, C.byref(ctx)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ecm = np.linspace(1.18, 2.0, n); p = C.c_void_p()
st = np.ascontiguousarray([[0, 0, 0], [1, 1, n - 1]], dtype=np.int32)
L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), 0.827, len(st), L.iptr(st), None, C.byref(p)))
L.check(lib.isr_assemble(p, None, None))
err = np.ones(n)
for _ in range(3): L.check(lib.isr_factor(p, L.dptr(err)))
NB = (n + 3) // 4
RW = int(os.environ.get("ISR_GJ_RW", "8")); CT = int(os.environ.get("ISR_GJ_CTAS", "8"))
RC = RW * ((n + RW * CT - 1) // (RW * CT))     # rows per CTA
tr = torch.zeros((NB, 8), dtype=torch.int64, device="cuda")
cl = C.CDLL(L.LIB_PATH)
cl.isr_debug_gj_trace.argtypes = [C.c_void_p]
assert cl.isr_debug_gj_trace(tr.data_ptr()) == 0
L.check(lib.isr_factor(p, L.dptr(err)))
torch.cuda.synchronize()
t = tr.cpu().numpy()
print("block: owner[start->credits, ->elim, ->publish] | next owner [wait_begin, visible - publish(B), applied - visible] | publish(B+1)-publish(B)")
for B in range(NB - 1):
    o = t[B]
    same_cta = (4 * B) // RC == (4 * (B + 1)) // RC
    print(f"{B:3d} {'   ' if same_cta else 'xct'} owner {o[1]-o[0]:6d} {o[2]-o[1]:6d} {o[3]-o[2]:5d} | next {o[4]-o[3]:7d} {o[5]-o[3]:6d} {o[6]-o[5]:6d} | {t[B+1][3]-o[3]:7d}")
d = [t[B + 1][3] - t[B][3] for B in range(NB - 1) if (4 * B) // RC == (4 * (B + 1)) // RC]
print("median cycles per block within a CTA:", int(np.median(d)))
