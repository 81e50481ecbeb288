#!/usr/bin/env python3
"""bench.py -- headline benchmark of the ISRSolver hot path on B200.

Metric (BASELINE.json): toy SLE solves/s for the N = 200 chi2 test, with the matrix-assembly time beside
it.  One "step" is one complete chi2 test on one batch of synthetic input:
    assemble the N x N integral-operator matrix  ->  factor  ->  T toys: b_k = G^-1 v_k, chi2_k
    ->  (N > 1: one all_gather of the packed per-rank results)  ->  min/max + TH1F(128) histogram.
Workload: N = 200 energy points on [1.18, 2.00] GeV, threshold 0.827 GeV, interpolation
[[linear, 0, 0], [cubic, 1, N-1]] (dense G), efficiency 1, Breit-Wigner Born model, 5 % errors,
T = 2^20 toys PER GPU (weak scaling), toys generated on the host side of the API contract
(V = vcs + err * z) -- for the device-resident `value` they are produced once on the device before the
timed region; for `e2e` they sit in pinned host memory and are copied inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--toys T] [--n N]
For N > 1 launch through torch.distributed.run (one rank per GPU, NCCL).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_STDOUT_FD = 1


def emit(line: dict):
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


E_MIN, E_MAX, E_THR = 1.18, 2.00, 0.827
BW_M, BW_G, BW_S0 = 1.50, 0.25, 4.0
FP64_PEAK_TFLOPS = 37.1      # measured DMMA peak on this pool's B200 (profiles/fp64_peaks_r01.json)


def bw_born(e):
    """Relativistic Breit-Wigner x phase space: the synthetic Born cross section."""
    e = np.asarray(e, dtype=np.float64)
    ps = np.clip(1.0 - E_THR ** 2 / e ** 2, 0.0, None)
    return BW_S0 * BW_M ** 2 * BW_G ** 2 / ((e * e - BW_M ** 2) ** 2 + BW_M ** 2 * BW_G ** 2) * ps ** 1.5


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(rows), "reasons": sorted(reasons)}


# =====================================================================================================
# reference arm: the reference's own CPU algorithm (assembly: its compiled code; toy loop: the LAPACK-backed oracle port,
# which is faster than the reference's compiled loop on the Eigen stand-in -- sampled beside it as `compiled_reference`)
# =====================================================================================================
def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import isr_oracle as O
    n = args.n
    cores = os.cpu_count() or 1
    ecm = np.linspace(E_MIN, E_MAX, n)
    settings = [(False, 0, 0), (True, 1, n - 1)]
    asm = reference_assembly(n, ecm, settings, cores)
    G = asm.pop("G")
    bcs0 = bw_born(ecm); vcs = G @ bcs0; err = 0.05 * vcs
    per_step = args.ref_toys
    V = O.draw_toys(vcs, err, per_step * (args.steps + args.warmup))
    for w in range(args.warmup):
        loop_on_all_cores(O, G, V[w * per_step:(w + 1) * per_step], bcs0, err, cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        b = (args.warmup + k) * per_step
        loop_on_all_cores(O, G, V[b:b + per_step], bcs0, err, cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    compiled = compiled_reference_sample(n, ecm, settings, G, vcs, err, bcs0, V[:8])
    sample = (f"{per_step} toys per step of the faithful per-toy loop (COD solve + G^T W G + inverse every toy, "
              f"src/Chi2Test.cpp:38-60) at N={n}, toys dealt to {cores} threads with ONE BLAS thread each (fixed, so that the "
              f"figure is stable from run to run).  Timed: the NumPy/LAPACK port of the loop, the FASTER of the two CPU forms "
              f"available ({compiled['note']}); assembly: {asm['sample']}")
    line = {"impl": "reference", "metric": "toy SLE solves/s (N=200 chi2 test)", "value": value, "unit": "toys/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": "toys/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "toys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "assembly": asm, "compiled_reference": compiled}
    emit(line)


def compiled_reference_sample(n, ecm, settings, G, vcs, err, bcs0, V):
    """The reference's OWN chi2TestModel (src/Chi2Test.cpp compiled into oracle/_ref against the dense Eigen stand-in of
    oracle/ref_shim) on a few toys, one thread: reported beside the port so that the choice of baseline is visible.  The
    stand-in's unblocked decompositions are slower than LAPACK, so the port is the conservative (faster) baseline."""
    from oracle import isr_ref as R
    if not R.available():
        return {"available": False, "note": "oracle/_ref not present on this box"}
    S = R.RefSolver("sle", ecm, vcs, err, E_THR, settings=settings)
    S.inject_matrix(G)
    t0 = time.perf_counter()
    S.chi2_test_model(V, vcs, bcs0)
    ms = (time.perf_counter() - t0) / (len(V) + 1) * 1e3          # chi2TestModel solves once before the loop
    return {"available": True, "ms_per_toy_1core": ms, "toys": len(V),
            "note": f"the reference's compiled chi2TestModel on the Eigen stand-in: {ms:.1f} ms per toy on one core"}


def loop_on_all_cores(O, G, V, bcs0, err, cores):
    """The reference's per-toy loop (oracle port) over a batch of toys: the toys are dealt to `cores` threads, each running
    the loop on its slice with a single BLAS thread (LAPACK releases the GIL).  Round 1 left the BLAS thread count to
    chance and the figure swung 7x between boxes."""
    from concurrent.futures import ThreadPoolExecutor
    from threadpoolctl import threadpool_limits
    parts = [V[q::cores] for q in range(cores) if len(V[q::cores])]
    with threadpool_limits(limits=1):
        with ThreadPoolExecutor(max_workers=len(parts)) as ex:
            res = list(ex.map(lambda part: O.chi2_test_loop(G, part, bcs0, err)[0], parts))
    return res


def reference_assembly(n, ecm, settings, cores):
    """Matrix assembly on the host cores: the reference's OWN compiled evalKuraevFadinBasisIntegral
    (oracle/_ref/libisr_ref.so) when that prebuilt library travelled with the repo, else the oracle port.
    Full matrix on all cores (the reference itself is single-threaded: 1-core time = cores x this, also
    sampled directly on three rows)."""
    from oracle import isr_oracle as O
    from oracle import isr_ref as R
    rows = (0, n // 2, n - 1)
    if R.available():
        ri = R.RefInterp(ecm, E_THR, settings)
        t0 = time.perf_counter(); ri.eval_eq_matrix(rows=rows, nthreads=1); row_s = (time.perf_counter() - t0) / len(rows)
        t0 = time.perf_counter(); G = ri.eval_eq_matrix(nthreads=cores); full_s = time.perf_counter() - t0
        kind = "reference"
    else:
        oi = O.Interp(ecm, E_THR, settings)
        t0 = time.perf_counter(); oi.eval_rows(rows, mode="faithful"); row_s = (time.perf_counter() - t0) / len(rows)
        t0 = time.perf_counter(); G = oi.eval_eq_matrix(mode="faithful", nthreads=cores); full_s = time.perf_counter() - t0
        kind = "port"
    return {"G": G, "ms": full_s * 1e3, "unit": "ms per N x N matrix", "cores": cores, "kind": kind,
            "ms_1core_from_rows": row_s * n * 1e3,
            "sample": (f"full {n} x {n} matrix, adaptive GK21/GK61 at 1e-12 per entry (src/Integration.cpp:19-81), "
                       f"{cores} threads: {full_s * 1e3:.0f} ms; rows {rows} on 1 thread: {row_s * 1e3:.0f} ms/row "
                       f"=> {row_s * n:.1f} s per matrix single-threaded, as the reference runs it")}


def workload_config(args, world):
    return {"workload": f"isrsolver-SLE-chi2-test: N={args.n} points, interp [[linear,0,0],[cubic,1,N-1]], eps=1, "
                        f"{args.toys} toys per GPU, full matrix assembly + factor + toy batch + histogram per step",
            "n_points": args.n, "toys_per_gpu": args.toys, "toys_total": args.toys * world,
            "l2_policy": f"inputs larger than L2 (V = {args.toys * args.n * 8 / 2**20:.0f} MiB per GPU streams through the 126 MB L2 every step)",
            "parallelism": f"toys sharded over {world} GPU(s), matrix replicated; per step, inside the C-ABI call: ncclAllReduce(MIN) of "
                           f"(min, -max) and ncclAllReduce(SUM) of ONE packed FP64 buffer (130 cells | sum_bcs | ratio stats | count | flag)"}


# =====================================================================================================
# B200 arm
# =====================================================================================================
class Tester:
    """isr_chi2_test (include/isr_b200.h) with its argument block: the whole chi2 test -- assemble, factor, toys, range,
    TH1F cells and, with a communicator, the two all-reduces -- in ONE stream-ordered C-ABI call."""

    def __init__(self, L, lib, prob, n, err, bcs0, vcs, comm, nbins=128, ratio=False, sum_bcs=True):
        self.L, self.lib, self.prob, self.n = L, lib, prob, n
        self.err, self.bcs0, self.vcs = (np.ascontiguousarray(x, dtype=np.float64) for x in (err, bcs0, vcs))
        self.counts = np.zeros(nbins + 2, dtype=np.int64)
        self.lohi, self.sumb, self.rstat = np.zeros(2), np.zeros(n), np.zeros(3 * n)
        self.total = C.c_int64()
        a = L.Chi2TestArgs()
        a.flags = L.ISR_TEST_ASSEMBLE | L.ISR_TEST_FACTOR | (L.ISR_TEST_RATIO if ratio else 0)
        a.vcs_err, a.bcs0, a.vcs = L.dptr(self.err), L.dptr(self.bcs0), L.dptr(self.vcs)
        a.nbins, a.comm = nbins, comm
        a.chi2_lo_hi_out = L.dptr(self.lohi)
        a.chi2_counts_out = self.counts.ctypes.data_as(C.POINTER(C.c_int64))
        if sum_bcs:
            a.sum_bcs_out = L.dptr(self.sumb)
        if ratio:
            a.ratio_stats_out = L.dptr(self.rstat)
        a.total_toys_out = C.pointer(self.total)
        self.a = a
        self.ms = np.zeros(4)

    def run(self, T, V_dev=None, V_host=None, chi2_host=None, seed=0, first_toy=0):
        a = self.a
        a.n_toys, a.V_dev, a.V, a.chi2_out = T, V_dev, V_host, chi2_host
        a.seed, a.first_toy = seed, first_toy
        self.L.check(self.lib.isr_chi2_test(self.prob, C.byref(a)))
        return self.counts

    def phases(self):
        self.L.check(self.lib.isr_chi2_test_timing(self.prob, self.L.dptr(self.ms)))
        return self.ms.copy()


def make_comm(L, lib, ctx, rank, world, torch, dist, dev):
    """One communicator of the C ABI per rank: rank 0's NCCL unique id travels over torch.distributed (plumbing); every
    collective of the measured path is then issued by libisr_b200 itself, in stream."""
    if world == 1:
        return None
    idbuf = torch.zeros(L.ISR_COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_ubyte * L.ISR_COMM_ID_BYTES)()
        L.check(lib.isr_comm_unique_id(raw))
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
    idbuf = idbuf.to(dev)
    dist.broadcast(idbuf, 0)
    raw = (C.c_ubyte * L.ISR_COMM_ID_BYTES)(*idbuf.cpu().tolist())
    comm = C.c_void_p()
    L.check(lib.isr_comm_init_rank(ctx, world, rank, raw, C.byref(comm)))
    return comm


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from isrsolver_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = L.load()
    ctx = C.c_void_p()
    L.check(lib.isr_ctx_create(local_rank, C.byref(ctx)))
    node = C.c_int(-1)
    L.check(lib.isr_ctx_bind_host_thread(ctx, C.byref(node)))     # this rank's host thread next to its GPU (NUMA)
    stream = torch.cuda.ExternalStream(lib.isr_ctx_stream(ctx), device=dev)
    comm = make_comm(L, lib, ctx, rank, world, torch, dist, dev)

    n, T = args.n, args.toys
    ecm = np.linspace(E_MIN, E_MAX, n)
    st = np.ascontiguousarray([[0, 0, 0], [1, 1, n - 1]], dtype=np.int32)
    prob = C.c_void_p()
    L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), E_THR, len(st), L.iptr(st), None, C.byref(prob)))
    G = np.empty((n, n))
    L.check(lib.isr_assemble(prob, None, L.dptr(G)))          # also the first-touch allocation of all workspaces
    G = G.T.copy()
    bcs0 = bw_born(ecm)
    vcs = G @ bcs0
    err = 0.05 * vcs

    # toys: per-rank stream of the documented generator shape (V = vcs + err * z); device copy for `value`,
    # page-locked host copy (isr_host_alloc: NUMA-local to this rank's GPU) for `e2e`
    gen = torch.Generator(device=dev); gen.manual_seed(20211103 + rank)
    z = torch.randn((T, n), dtype=torch.float64, device=dev, generator=gen)
    Vd = torch.from_numpy(vcs).to(dev)[None, :] + torch.from_numpy(err).to(dev)[None, :] * z
    del z
    vh_ptr, c2_ptr = C.c_void_p(), C.c_void_p()
    L.check(lib.isr_host_alloc(ctx, T * n * 8, C.byref(vh_ptr)))
    L.check(lib.isr_host_alloc(ctx, T * 8, C.byref(c2_ptr)))
    Vh = np.ctypeslib.as_array(C.cast(vh_ptr, C.POINTER(C.c_double)), shape=(T, n))
    chi2_host = np.ctypeslib.as_array(C.cast(c2_ptr, C.POINTER(C.c_double)), shape=(T,))
    torch.from_numpy(Vh).copy_(Vd)
    torch.cuda.synchronize()
    test = Tester(L, lib, prob, n, err, bcs0, vcs, comm)
    Vh_p, c2_p = C.cast(vh_ptr, C.POINTER(C.c_double)), C.cast(c2_ptr, C.POINTER(C.c_double))

    ev = lambda: torch.cuda.Event(enable_timing=True)
    phase_ms = []

    def step_device(record):
        """One chi2 test with V resident in HBM."""
        tot = test.run(T, V_dev=Vd.data_ptr())
        if record:
            phase_ms.append(test.phases())
        return tot

    def step_e2e():
        """The same test with HOST buffers: H2D of V and D2H of chi2 inside the call."""
        return test.run(T, V_host=Vh_p, chi2_host=c2_p)

    def step_campaign():
        """The same test asked for BY COUNT, as the reference's chi2Test(n, ...) is: toys drawn on the device
        (Philox, csrc/rng.cu); only vcs / err / bcs0 go in, the TH1F cells and sums come out."""
        return test.run(T, seed=20211103, first_toy=rank * T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, record=False):
        barrier()
        e0, e1 = ev(), ev()
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(steps):
            fn(record) if record is not None else fn()
        with torch.cuda.stream(stream):
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        hist = step_device(False).copy()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    launches0 = lib.isr_launch_count()
    ms_total = timed(step_device, args.steps, record=True)
    launches = lib.isr_launch_count() - launches0
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e = timed(lambda: step_e2e(), args.steps, record=None)
    hist_c = step_campaign().copy()
    ms_rng = timed(lambda: step_campaign(), args.steps, record=None)
    clocks = sampler.finish() if sampler else None      # sampled every 25 ms across the three timed regions (value, e2e, campaign)
    assert int(hist_c.sum()) == T * world and test.total.value == T * world
    assert int(hist.sum()) == T * world, "histogram lost entries"
    strong = strong_scaling(L, lib, ctx, stream, comm, torch, dist, dev, rank, world)
    c4 = tikhonov_c4(lib, ctx, torch, comm, rank, world)
    if rank != 0:
        return
    ms_step = ms_total / args.steps
    value = T * world / (ms_step * 1e-3)
    ph = np.mean(np.array(phase_ms), axis=0)
    asm_ms, fac_ms, toy_ms, tail_ms = (float(x) for x in ph)
    flops = 4.0 * n * n * T          # 2N^2 for b = G^-1 v, 2N^2 for |W^(1/2) G (b - b0)|^2
    achieved = flops / (toy_ms * 1e-3) * 1e-12
    npan, nnode = C.c_int64(), C.c_int64()
    L.check(lib.isr_assemble_stats(prob, C.byref(npan), C.byref(nnode)))
    line = {
        "metric": "toy SLE solves/s (N=200 chi2 test)", "value": value, "unit": "toys/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "assembly_ms": asm_ms, "factor_ms": fac_ms, "toy_batch_ms": toy_ms, "tail_ms": tail_ms,
        "assembly": {"panels": npan.value, "nodes": nnode.value, "ms": asm_ms},
        "roofline": {"bound": "tensor", "kernel": "toy_fused_kernel, tile " + (lib.isr_toy_tile_name(n, 1) or b"?").decode() +
                               " (one launch: TMA box loads -> DMMA m8n8k4 f64, pass 1 b = S v and pass 2 chi2 = |R(b-b0)|^2)",
                     "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                     # dram__bytes_read + dram__bytes_write of the fused launch from the ncu --set full capture of THIS
                     # command in profiles/ncu_r02_toy_dense.txt (1679.86 + 202.05 MB for 1048576 toys at N=200), scaled to T
                     "traffic": (1881.91e6 / 1048576) * T if n == 200 else None,
                     "algorithmic_bytes": (8 * n + 8) * T, "flops_per_toy": 4 * n * n,
                     "timing": "CUDA events recorded by isr_chi2_test on its own stream around the fused launch (isr_chi2_test_timing), mean over the timed steps",
                     "peak_source": "FP64 DMMA probe tools/fp64_peaks.cu -> profiles/fp64_peaks_r01.json "
                                    "(MEASURED_PEAKS.json carries no FP64 figure; bf16 peak does not apply to an FP64 kernel)"},
        "e2e": {"value": T * world / (ms_e2e / args.steps * 1e-3), "unit": "toys/s",
                "h2d_bytes_per_step": T * n * 8, "d2h_bytes_per_step": T * 8 + (130 + 4 * n + 4) * 8, "ms_per_step": ms_e2e / args.steps,
                "h2d_gb_per_s_per_gpu": T * n * 8 / (ms_e2e / args.steps * 1e-3) * 1e-9, "numa_node_of_rank0": node.value,
                "note": "isr_chi2_test with host buffers from isr_host_alloc (page-locked, NUMA-local to the GPU): PCIe-bound"},
        "campaign_by_count": {"value": T * world / (ms_rng / args.steps * 1e-3), "unit": "toys/s", "ms_per_step": ms_rng / args.steps,
                              "h2d_bytes_per_step": 3 * n * 8, "d2h_bytes_per_step": (130 + 4 * n + 4) * 8,
                              "note": "same chi2 test with toys drawn on the device (Philox4x32-10 + Box-Muller, as the reference's "
                                      "chi2Test(n) draws them internally); NOT the e2e figure -- its toy vectors never cross PCIe"},
        "gpu_launches": int(launches), "clocks": clocks,
        "collectives": "none (1 GPU)" if world == 1 else "per step, inside isr_chi2_test: all-reduce(MIN) of {min, -max} (2 doubles) + all-reduce(SUM) of one "
                       "packed FP64 buffer [130 cells | sum b (N) | ratio stats (3N) | toys | flag], via " +
                       ("NVLink peer-memory mailboxes fused into the range / cells kernels" if lib.isr_comm_uses_peer_memory(comm) else "ncclAllReduce on the context stream"),
        "target_strong": strong,
    }
    line["assembly_other_configs"] = other_assemblies(lib, ctx, stream, torch)
    line["chi2_test_other_configs"] = other_chi2_tests(lib, ctx, stream, torch, dev)
    line["tikhonov_c4"] = c4
    line["cond_spread_scan"] = cond_scan(lib, ctx)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(n, G, bcs0, vcs, err)
    emit(line)


def _problem_for(L, lib, ctx, n, cubic):
    ecm = np.linspace(E_MIN, E_MAX, n)
    st = np.ascontiguousarray([[0, 0, 0], [1, 1, n - 1]] if cubic else [[0, 0, n - 1]], dtype=np.int32)
    p = C.c_void_p()
    L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), E_THR, len(st), L.iptr(st), None, C.byref(p)))
    G = np.empty((n, n))
    L.check(lib.isr_assemble(p, None, L.dptr(G)))
    bcs0 = bw_born(ecm); vcs = G.T @ bcs0; err = 0.05 * vcs
    return p, bcs0, vcs, err


def strong_scaling(L, lib, ctx, stream, comm, torch, dist, dev, rank, world):
    """STRONG scaling, every rank: the north-star target (N = 200 cubic, 2^20 toys IN TOTAL: 2^20 / world per GPU) and
    BASELINE config C5 (N = 500 linear, 2^20 toys in total, chi2 + ratio statistics), each as one isr_chi2_test per
    rank -- assembly + factorisation + toys + the two all-reduces + TH1F cells -- with V resident.  Time = CUDA events
    around 5 tests after 2 warm-ups, max over ranks.  frac_of_fp64_peak = algorithmic flop of ALL toys / (world x time x
    37.1 TFLOP/s); C5's operators are lower triangular, so its algorithmic count is 2 N^2 per toy."""
    out = {}
    total = 1 << 20
    for name, n, cubic, ratio in (("target_n200_cubic_1M_toys_total", 200, True, False),
                                  ("c5_n500_linear_1M_toys_total_ratio", 500, False, True)):
        Tr = total // world
        p, bcs0, vcs, err = _problem_for(L, lib, ctx, n, cubic)
        gen = torch.Generator(device=dev); gen.manual_seed(11 + rank)
        Vd = torch.from_numpy(vcs).to(dev)[None, :] + torch.from_numpy(err).to(dev)[None, :] * torch.randn((Tr, n), dtype=torch.float64, device=dev, generator=gen)
        t = Tester(L, lib, p, n, err, bcs0, vcs, comm, ratio=ratio, sum_bcs=ratio)
        for _ in range(2):
            t.run(Tr, V_dev=Vd.data_ptr())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(reps):
            cnt = t.run(Tr, V_dev=Vd.data_ptr())
        with torch.cuda.stream(stream):
            e1.record(stream)
        torch.cuda.synchronize()
        phases = [t.phases()]        # (of the last test; read outside the timed loop)
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        assert int(cnt.sum()) == total and t.total.value == total
        ph = np.mean(np.array(phases), axis=0)
        flop_per_toy = (4 if cubic else 2) * n * n
        out[name] = {"ms_per_test": ms, "toys_total": total, "toys_per_gpu": Tr, "n_points": n, "toys_per_s": total / (ms * 1e-3),
                     "frac_of_fp64_peak": flop_per_toy * total / world / (ms * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS,
                     "flop_per_toy": flop_per_toy, "assembly_ms": float(ph[0]), "factor_ms": float(ph[1]), "toys_ms": float(ph[2]),
                     "tail_ms": float(ph[3]), "toy_schedule": "dense" if cubic else "triangular",
                     "outputs": "TH1F(128) cells and range" + (", ratio statistics, sum of solutions" if ratio else " (chi2TestModel: no column sums)")}
        del Vd
        lib.isr_problem_destroy(p)
    return out


def other_assemblies(lib, ctx, stream, torch):
    """Matrix-assembly ms (device-resident, CUDA events, mean of 10 after 2 warm-ups) for the other BASELINE.json
    configurations: C1 N=50 linear, C2 N=100 cubic, C3 N=200 with a 512-bin efficiency histogram per point
    (ISRSolverSLE2, 'matrix-assembly bound'), C5 N=500 linear."""
    from isrsolver_b200 import _lib as L
    out = {}
    for name, n, cubic, nbins in (("c1_n50_linear", 50, False, 0), ("c2_n100_cubic", 100, True, 0),
                                  ("c3_n200_eps512xN", 200, False, 512), ("c5_n500_linear", 500, False, 0)):
        ecm = np.linspace(E_MIN, E_MAX, n)
        st = np.ascontiguousarray([[0, 0, 0], [1, 1, n - 1]] if cubic else [[0, 0, n - 1]], dtype=np.int32)
        desc = None
        if nbins:
            xc = (np.arange(nbins) + 0.5) / nbins
            vals = np.zeros((n, nbins + 2))
            vals[:, 1:-1] = 0.25 + 0.5 * np.exp(-40.0 * xc)[None, :] * (1.0 - 0.1 * (ecm[:, None] - 1.5))
            desc = L.EpsDesc()
            desc.kind, desc.nx, desc.xlo, desc.xhi, desc.values = 1, nbins, 0.0, 1.0, L.dptr(vals)
        p = C.c_void_p()
        L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), E_THR, len(st), L.iptr(st), C.byref(desc) if desc else None, C.byref(p)))
        for _ in range(2):
            L.check(lib.isr_assemble(p, None, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(10):
                L.check(lib.isr_assemble(p, None, None))
            e1.record(stream)
        e1.synchronize()
        npan, nnode = C.c_int64(), C.c_int64()
        L.check(lib.isr_assemble_stats(p, C.byref(npan), C.byref(nnode)))
        out[name] = {"ms": e0.elapsed_time(e1) / 10, "panels": npan.value, "nodes": nnode.value}
        lib.isr_problem_destroy(p)
    return out


def other_chi2_tests(lib, ctx, stream, torch, dev):
    """The same step (one isr_chi2_test: assemble + factor + toys + min/max + TH1F) at the other BASELINE.json chi2-test
    sizes, V resident, this rank only: C2 = N 100, cubic, 100k toys (configs[1]); C5's per-GPU share = N 500, linear, 2^17
    toys with the ratio statistics on top -- lower-triangular operators, so the toy kernel runs its triangular schedule
    (about 2 N^2 flop per toy executed); and the north-star target's per-GPU share: N 200, cubic, 2^17 toys (1M toys on
    8 GPUs), with the fraction of the FP64 roofline the WHOLE test reaches (4 N^2 T flop over the time of assembly +
    factorisation + toys + histogram).  CUDA events, mean of 5 after 2 warm-ups."""
    from isrsolver_b200 import _lib as L
    out = {}
    for name, n, cubic, T, ratio in (("c2_n100_cubic_100k_toys", 100, True, 100000, False),
                                     ("c5_n500_linear_131072_toys_ratio", 500, False, 1 << 17, True),
                                     ("target_n200_cubic_131072_toys", 200, True, 1 << 17, False)):
        p, bcs0, vcs, err = _problem_for(L, lib, ctx, n, cubic)
        gen = torch.Generator(device=dev); gen.manual_seed(7)
        Vd = torch.from_numpy(vcs).to(dev)[None, :] + torch.from_numpy(err).to(dev)[None, :] * torch.randn((T, n), dtype=torch.float64, device=dev, generator=gen)
        t = Tester(L, lib, p, n, err, bcs0, vcs, None, ratio=ratio, sum_bcs=ratio)      # chi2TestModel needs no column sums
        torch.cuda.synchronize()
        for _ in range(2):
            t.run(T, V_dev=Vd.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(5):
            h = t.run(T, V_dev=Vd.data_ptr())
        with torch.cuda.stream(stream):
            e1.record(stream)
        e1.synchronize()
        phases = [t.phases()]        # (of the last test; read outside the timed loop)
        assert int(h.sum()) == T
        ms = e0.elapsed_time(e1) / 5
        ph = np.mean(np.array(phases), axis=0)
        flop_per_toy = (4 if cubic else 2) * n * n
        out[name] = {"ms_per_test": ms, "toys_per_s": T / (ms * 1e-3), "n_points": n, "toys": T,
                     "assembly_ms": float(ph[0]), "factor_ms": float(ph[1]), "toys_ms": float(ph[2]), "tail_ms": float(ph[3]),
                     "toy_schedule": "dense (4 N^2 flop/toy)" if cubic else "triangular (2 N^2 flop/toy algorithmic)",
                     "whole_test_frac_of_fp64_peak": flop_per_toy * T / (ms * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS}
        lib.isr_problem_destroy(p)
    return out


def tikhonov_c4(lib, ctx, torch, comm=None, rank=0, world=1):
    """BASELINE config C4, every rank: ISRSolverTikhonov, N = 200, cubic, 2 MeV energy spread, derivative regulariser;
    the generalised eigen-system once (isr_tikhonov_prepare), then the chi2 test of 10^4 toys for 1024 lambdas with the
    TOYS sharded over the ranks (isr_tikhonov_chi2_scan: 10^4 / world toys per GPU from page-locked host memory, H2D and
    D2H inside the call, per-lambda ranges by one all-reduce(MIN), cells and sums by one all-reduce(SUM)).  Wall clock
    around synchronous calls, mean of 5 after 2 warm-ups, max over ranks."""
    import torch.distributed as dist
    from isrsolver_b200 import _lib as L
    n, T, Lm = 200, 10000, 1024
    ecm = np.linspace(E_MIN, E_MAX, n)
    st = np.ascontiguousarray([[0, 0, 0], [1, 1, n - 1]], dtype=np.int32)
    p = C.c_void_p()
    L.check(lib.isr_problem_create(ctx, n, L.dptr(ecm), E_THR, len(st), L.iptr(st), None, C.byref(p)))
    sig = np.full(n, 0.002)
    G = np.empty((n, n))
    L.check(lib.isr_assemble(p, L.dptr(sig), L.dptr(G)))
    bcs0 = bw_born(ecm); vcs = G.T @ bcs0; err = 0.05 * vcs
    lam = 1e-9 * (1e9) ** (np.arange(Lm) / (Lm - 1))          # src/isrsolver-LCurve-plot.cpp:127-132
    rng = np.random.default_rng(20211103)
    Vall = vcs[None, :] + err[None, :] * rng.standard_normal((T, n))
    t0, t1 = rank * T // world, (rank + 1) * T // world       # this rank's toys
    vptr = C.c_void_p()
    L.check(lib.isr_host_alloc(ctx, max(t1 - t0, 1) * n * 8, C.byref(vptr)))
    Vp = np.ctypeslib.as_array(C.cast(vptr, C.POINTER(C.c_double)), shape=(max(t1 - t0, 1), n))
    Vp[:t1 - t0] = Vall[t0:t1]
    stats = np.empty((Lm, 3))
    cptr = C.c_void_p()
    L.check(lib.isr_host_alloc(ctx, Lm * 130 * 4, C.byref(cptr)))        # the caller's cell buffer, page-locked like the toys
    counts = np.ctypeslib.as_array(C.cast(cptr, C.POINTER(C.c_uint32)), shape=(Lm, 130))
    total = C.c_int64()

    def wall(fn, reps=5):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t_0) / reps * 1e3
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms
    prep = wall(lambda: L.check(lib.isr_tikhonov_prepare(p, L.dptr(err), 1)))
    a = L.TikhonovScanArgs()
    a.flags, a.deriv_reg, a.n_lambda, a.lambda_, a.n_toys = 0, 1, Lm, L.dptr(lam), t1 - t0
    a.V, a.bcs0, a.nbins, a.comm = C.cast(vptr, C.POINTER(C.c_double)), L.dptr(bcs0), 128, comm
    a.stats_out, a.counts_out, a.total_toys_out = L.dptr(stats), L.uptr(counts), C.pointer(total)
    full = wall(lambda: L.check(lib.isr_tikhonov_chi2_scan(p, C.byref(a))))
    assert int(counts.sum()) == Lm * T and total.value == T
    a.flags, a.vcs_err = L.ISR_TEST_FACTOR, L.dptr(err)
    both = wall(lambda: L.check(lib.isr_tikhonov_chi2_scan(p, C.byref(a))))
    out = {"n_points": n, "lambdas": Lm, "toys_per_lambda": T, "toys_per_gpu": t1 - t0, "prepare_ms": prep,
           "chi2_test_1024_lambdas_ms": full, "pairs_per_s": Lm * T / (full * 1e-3),
           "eigen_system_plus_chi2_test_ms": both,
           "note": "toys sharded over the GPUs (isr_tikhonov_chi2_scan); the eigen-system (cuSOLVER sygvd) is replicated; "
                   "eigen_system_plus_chi2_test_ms = both in one call, the H2D of the toys under the eigen-system"}
    if rank == 0:
        lc = [np.empty(Lm) for _ in range(4)]
        t_0 = time.perf_counter()
        for _ in range(5):
            L.check(lib.isr_tikhonov_scan(p, Lm, L.dptr(lam), L.dptr(vcs), L.dptr(lc[0]), L.dptr(lc[1]), L.dptr(lc[2]), L.dptr(lc[3]), None))
        out["lcurve_1024_lambdas_ms"] = (time.perf_counter() - t_0) / 5 * 1e3
    lib.isr_host_free(vptr)
    lib.isr_host_free(cptr)
    lib.isr_problem_destroy(p)
    return out


def cond_scan(lib, ctx):
    """Condition number of G_spread(sigma) G for 100 beam-energy spreads at N = 200 (src/isrsolver-condnum-test.cpp:88-140,
    notebooks/condnum_enspread.ipynb): one assembly, then batched spread matrices, products and one-sided Jacobi singular
    values (one CTA per setting).  Wall clock of the synchronous call, best of 3."""
    from isrsolver_b200 import _lib as L
    n, K = 200, 100
    p, _, _, _ = _problem_for(L, lib, ctx, n, True)
    sig = np.ascontiguousarray(np.linspace(0.0, 0.005, K))
    cond = np.zeros(K)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        L.check(lib.isr_cond_spread_scan(p, K, L.dptr(sig), 0, L.dptr(cond), None))
        best = min(best, (time.perf_counter() - t0) * 1e3)
    lib.isr_problem_destroy(p)
    return {"n_points": n, "settings": K, "ms": best, "cond_first": float(cond[0]), "cond_last": float(cond[-1]),
            "note": "round 1: 1.1 s (100 x cuSOLVER gesvd on 8 streams)"}


def cpu_baseline(n, G, bcs0, vcs, err):
    """The oracle (port of the reference's per-toy loop) on the box's host cores, bounded sample."""
    from oracle import isr_oracle as O
    cores = os.cpu_count() or 1
    V = O.draw_toys(vcs, err, 4 * cores)
    t0 = time.perf_counter(); loop_on_all_cores(O, G, V, bcs0, err, cores); per = (time.perf_counter() - t0) / len(V)
    m = int(max(64, min(4096, 12.0 / per)))
    V = O.draw_toys(vcs, err, m)
    t0 = time.perf_counter(); loop_on_all_cores(O, G, V, bcs0, err, cores); dt = time.perf_counter() - t0
    Vf = O.draw_toys(vcs, err, 100000)
    t1 = time.perf_counter(); O.chi2_test_fair(G, Vf, bcs0, err); dtf = time.perf_counter() - t1
    asm = reference_assembly(n, np.linspace(E_MIN, E_MAX, n), [(False, 0, 0), (True, 1, n - 1)], os.cpu_count() or 1)
    asm.pop("G")
    return {"value": m / dt, "unit": "toys/s", "cores": os.cpu_count() or 1, "kind": "port", "assembly": asm,
            "sample": f"{m} toys of the faithful per-toy loop (COD + G^T W G + inverse per toy, src/Chi2Test.cpp:38-60) at N={n}, NumPy/LAPACK, toys dealt to all host cores with one BLAS thread each",
            "fair_value": 100000 / dtf,
            "fair_sample": "100000 toys with the redundancy removed (factor once, two GEMMs; BASELINE.md B-fair), NumPy/OpenBLAS on all host cores"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--toys", type=int, default=1 << 20, help="toys per GPU")
    ap.add_argument("--ref-toys", type=int, default=256, help="toys per step of the reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout must carry exactly ONE JSON line: libraries that print there (NCCL's version banner does) are sent to
    # stderr for the whole run, and the line is written to the saved descriptor at the end
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
