"""bench.py -- PCG iterations/s on the 5-point Laplacian, n = 4096^2 = 16.8 M
(BASELINE.json metric; SURVEY.md section 8d, configuration C3).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one call of the hot path on one batch of synthetic input: a PCG
solve of exactly ITERS_PER_STEP iterations (tau = 0, failOnMaxiter = False) of
the un-preconditioned system -FDLaplacian2D(0,1,4096) x = 1.

* ``value``    iterations/s with A, b resident in HBM, CUDA events around K steps
* ``e2e``      iterations/s through the public API ``PCGSolver.solve(A, b)`` with
               HOST operands (scipy CSR + numpy vector in pinned memory): every
               step uploads A and b and downloads x and the residual history
* ``roofline`` dominant kernel = the persistent PCG kernel (one launch per step):
               algorithmic bytes 32 n + 200 (12 nnz + 4 (n+1) + 88 n) / CUDA-event time
               of the step; ``spmv_roofline``: the stand-alone SpMV fused with p.Ap,
               12 nnz + 4 (n+1) + 16 n per launch / mean launch time (events)
* ``iter_roofline``  the whole iteration: 12 nnz + 4 (n+1) + 88 n bytes / time
* ``ic_pcg``   configs[2] as named (PCG + incomplete Cholesky) at m = 1024, the largest
               size whose SuperLU setup fits a bench run
* ``parity``   (every N) small systems solved in-process through the same path right before the
               timed region and compared with scipy / the oracle: SpMV bit-exact, PCG residual
               histories (2-D, 3-D) to 1e-10, iteration counts, second solve on the same plan
               bit-identical to the first -- at N > 1 this exercises the fused peer-memory path
* ``c4``       (every N) configs[3]: 3-D 7-point Laplacian 512^3 (134 M unknowns), assembled on the
               device, same 200-iteration step; value / roofline fields as the main line
* ``cpu_baseline``   the oracle port (same numpy/scipy calls as the reference)
                     timed on this host on a bounded sample of the workload

``--impl reference`` times the reference's CPU implementation of the path (the
oracle port: the reference is pure Python over numpy/scipy, so there is nothing
to compile under oracle/_ref and the port executes the identical library calls).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M_GRID = int(os.environ.get('PSB_BENCH_M', '4096'))      # 4096^2 = 16 777 216 unknowns
ITERS_PER_STEP = int(os.environ.get('PSB_BENCH_ITERS', '200'))
METRIC = 'pcg_iterations_per_second'
UNIT = 'iter/s'


def workload_name(m):
    return ('2-D 5-point Laplacian m=%d (n=%d), un-preconditioned PCG, b=1, '
            '%d iterations per step' % (m, m * m, ITERS_PER_STEP))


def c4_name(m):
    return ('3-D 7-point Laplacian m=%d (n=%d), un-preconditioned PCG, b=1, %d iterations per step'
            % (m, m ** 3, ITERS_PER_STEP))


def bench_config(m, world):
    """The ``config`` object of the JSON line -- the same dict for the B200 arm and the
    reference arm at a given N (what is specific to an arm lives outside ``config``)."""
    n, nnz = m * m, 5 * m * m - 4 * m
    iter_bytes = 12 * nnz + 4 * (n + world) + 88 * n
    return {'workload': workload_name(m), 'n': n, 'nnz': nnz, 'iters_per_step': ITERS_PER_STEP,
            'partition': 'block rows over %d rank(s)' % world,
            'l2': 'inputs larger than L2: %.2f GB touched per iteration per rank vs 126 MB L2'
                  % (iter_bytes / world / 1e9)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed
    ncu --set full summaries under profiles/ (None when no capture names the kernel)."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_ncu_full.json'))):
        try:
            for d in json.load(open(path)):
                v = d.get('dram_traffic_bytes')
                if kernel_substr in d.get('kernel', '') and v and v == v:      # skip NaN captures
                    best = float(v)
        except Exception:
            pass
    return best


def build_problem(m, pinned=True):
    """A = -FDLaplacian2D(0,1,m) (examples/FDLaplacian2D.py semantics, stored
    column order kept), b = ones.  With ``pinned`` the CSR arrays and b live in
    page-locked host memory so the e2e leg copies from pinned memory."""
    from pysolvers_b200.problems import fd_laplacian_2d
    import scipy.sparse as sp
    A = fd_laplacian_2d(0.0, 1.0, m)
    A.data *= -1.0
    b = np.ones(A.shape[0])
    if pinned:
        import torch
        if torch.cuda.is_available():
            def pin(a):
                t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                return t.numpy()
            A = sp.csr_matrix((pin(A.data), pin(A.indices), pin(A.indptr)), shape=A.shape)
            b = pin(b)
    return A, b


def cpu_oracle_rate(A, b, iters, repeats=1):
    """iterations/s of the oracle port (= the reference's numpy/scipy calls)."""
    from oracle import krylov
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = krylov.pcg(A, b, maxiter=iters, tau=0.0, fail_on_maxiter=False)
        dt = time.perf_counter() - t0
        assert len(r['hist']) == iters
        best = dt if best is None else min(best, dt)
    return iters / best, best


def cpu_rates_both(A, b, iters):
    """(rate with every BLAS thread the host offers, rate with 1 BLAS thread, threads used):
    SURVEY.md section 8d asks for both -- scipy's csr_matvec is single-threaded either way, and
    OpenBLAS's threaded ddot is not always the faster one at these sizes.  The thread count is
    set explicitly so that it does not depend on how the process was launched (torchrun
    exports OMP_NUM_THREADS=1)."""
    ncpu = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        r, _ = cpu_oracle_rate(A, b, iters)
        return r, r, blas_threads()
    with threadpool_limits(limits=ncpu, user_api='blas'):
        used = blas_threads()
        r_all, _ = cpu_oracle_rate(A, b, iters)
    with threadpool_limits(limits=1, user_api='blas'):
        r_one, _ = cpu_oracle_rate(A, b, iters)
    return r_all, r_one, used


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        for lib in threadpool_info():
            if lib.get('user_api') == 'blas':
                return int(lib.get('num_threads', 1))
    except Exception:
        pass
    return 1


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on this host."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    A, b = build_problem(M_GRID, pinned=False)
    iters = int(os.environ.get('PSB_REF_ITERS_PER_STEP', '4'))   # bounded sample per step
    ncpu = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        limit_all = threadpool_limits(limits=ncpu, user_api='blas')     # same thread count at every N
    except Exception:
        limit_all = contextlib.nullcontext()
    with limit_all:
        cores = blas_threads()
        for _ in range(args.warmup):
            cpu_oracle_rate(A, b, 1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_oracle_rate(A, b, iters)
        dt = time.perf_counter() - t0
    value = args.steps * iters / dt
    try:
        with threadpool_limits(limits=1, user_api='blas'):
            one_rate, _ = cpu_oracle_rate(A, b, iters)
    except Exception:
        one_rate = None
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': bench_config(M_GRID, world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'value_1_blas_thread': one_rate,
                         'sample': '%d steps x %d PCG iterations of the full %d-row system (the config\'s '
                                   '%d-iteration step is sampled); scipy csr_matvec is single-threaded, '
                                   'OpenBLAS ddot pinned to %d thread(s) at every N; os.cpu_count()=%d'
                                   % (args.steps, iters, A.shape[0], ITERS_PER_STEP, cores, ncpu)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------
# parity block and the 512^3 side leg: run at every N (N = 1: single-GPU path)
# ---------------------------------------------------------------------------------
def parity_checks(comm, rank, world):
    """Small systems through the SAME path the timed region uses (N = 1: PCGSolver /
    psb_pcg_solve; N > 1: DistCSR + DistributedPCG, i.e. the fused peer-memory kernel),
    checked against scipy (SpMV, bit for bit) and the oracle (tests/dist_worker.py's checks,
    in-process).  Returns the ``parity`` object of the JSON line."""
    import torch
    import scipy.sparse as sp
    from oracle import krylov
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import PCG
    from pysolvers_b200.device import DeviceCSR
    from pysolvers_b200 import dist as pdist
    from pysolvers_b200.problems import fd_laplacian_2d, fd_laplacian_3d

    cases = [('lap2d_m96', -fd_laplacian_2d(0.0, 1.0, 96), 1e-8, 600, True),
             ('lap3d_m24', fd_laplacian_3d(0.0, 1.0, 24), 1e-9, 400, True),
             ('lap2d_m512_100it', -fd_laplacian_2d(0.0, 1.0, 512), 0.0, 100, False)]
    out = {'spmv_bitexact': True, 'second_solve_bitexact': True, 'hist_rel_err': 0.0,
           'solution_rel_err': 0.0, 'iters': [], 'iters_oracle': [], 'cases': [c[0] for c in cases]}
    for name, A, tau, maxiter, fail in cases:
        A = sp.csr_matrix(A)
        n = A.shape[0]
        b = np.ones(n)
        x = np.random.default_rng(3).standard_normal(n)
        ctl = CommonSolverArgs(maxiter=maxiter, tau=tau, failOnMaxiter=fail, showIters=False, showFinal=False)
        hist = []
        if world == 1:
            lo, hi = 0, n
            y = DeviceCSR(A).matvec(torch.from_numpy(x).cuda()).cpu().numpy()
            solver = PCG(ctl).makeSolver()
            solve = lambda: solver.solve(A, b)
        else:
            starts = pdist.row_starts(n, world)
            lo, hi = int(starts[rank]), int(starts[rank + 1])
            blk = A[lo:hi, :]
            D = pdist.DistCSR(comm, blk.indptr, blk.indices, blk.data, lo, hi, n)
            y = D.matvec(torch.from_numpy(x[lo:hi]).cuda()).cpu().numpy()
            solver = pdist.DistributedPCG(ctl)
            solve = lambda: solver.solve(D, b[lo:hi])
        solver.reportIter = lambda k, nr, nb: hist.append(nr)
        if not np.array_equal(y, (A @ x)[lo:hi]):
            out['spmv_bitexact'] = False
        with contextlib.redirect_stdout(io.StringIO()):
            st = solve()
            h1 = np.asarray(hist)
            st2 = solve()                      # epochs / ping-pong buffers / plan carry over
        if st2.iters() != st.iters() or not np.array_equal(st2.soln(), st.soln()):
            out['second_solve_bitexact'] = False
        ref = krylov.pcg(A, b, maxiter=maxiter, tau=tau, fail_on_maxiter=fail)
        k = min(len(h1), len(ref['hist']))
        rel = float(np.max(np.abs(h1[:k] - ref['hist'][:k]) / ref['hist'][:k])) if k else 1.0
        out['hist_rel_err'] = max(out['hist_rel_err'], rel)
        err = float(np.linalg.norm(st.soln() - ref['soln'][lo:hi]) / np.linalg.norm(ref['soln'][lo:hi]))
        out['solution_rel_err'] = max(out['solution_rel_err'], err)
        out['iters'].append(int(st.iters()))
        out['iters_oracle'].append(int(ref['iters']))
        if world > 1:
            D.close()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([out['hist_rel_err'], out['solution_rel_err'],
                          0.0 if out['spmv_bitexact'] else 1.0,
                          0.0 if out['second_solve_bitexact'] else 1.0], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out['hist_rel_err'], out['solution_rel_err'] = float(t[0]), float(t[1])
        out['spmv_bitexact'], out['second_solve_bitexact'] = bool(t[2] == 0), bool(t[3] == 0)
    out['ok'] = bool(out['spmv_bitexact'] and out['second_solve_bitexact'] and out['hist_rel_err'] <= 1e-10
                     and out['solution_rel_err'] <= 1e-8
                     and all(abs(a - b) <= 1 for a, b in zip(out['iters'], out['iters_oracle'])))
    out['tolerances'] = {'hist_rel_err': 1e-10, 'solution_rel_err': 1e-8, 'iters': 1}
    return out


def c4_leg(comm, rank, world, steps, warmup):
    """configs[3]: 3-D 7-point Laplacian m^3 (m = 512: 134 M unknowns), un-preconditioned PCG,
    block rows over the ranks, matrix assembled on the device (bit-identical to the host
    generator, tests/test_gpu_spmv.py).  Same step as the main line; device-resident timing."""
    import ctypes as C
    import torch
    from pysolvers_b200 import _native as nat
    from pysolvers_b200 import dist as pdist
    from pysolvers_b200.device import DeviceCSR, current_stream_ptr, ptr
    from pysolvers_b200.problems import device_fd_laplacian
    lib = nat.lib()
    m = int(os.environ.get('PSB_BENCH_M3', '512'))
    n = m ** 3
    iters = ITERS_PER_STEP
    starts = pdist.row_starts(n, world)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    n_loc = hi - lo
    dev = torch.device('cuda', torch.cuda.current_device())
    indptr, cols, data = device_fd_laplacian(3, 0.0, 1.0, m, row_lo=lo, row_hi=hi, raw=True)
    nnz_loc = int(data.numel())
    b_d = torch.ones(n_loc, dtype=torch.float64, device=dev)
    x_d = torch.empty(n_loc, dtype=torch.float64, device=dev)
    hist_d = torch.empty(iters, dtype=torch.float64, device=dev)
    res = nat.SolveResult()
    stream = current_stream_ptr()
    if world == 1:
        dA = DeviceCSR(indptr=indptr, indices=cols, data=data, shape=(n, n))
        wbytes = int(lib.psb_pcg_workspace_bytes(n, 0))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        mode = 'single GPU'

        def step():
            nat.check(lib.psb_pcg_solve(dA.handle, None, ptr(b_d), ptr(x_d), ptr(work), wbytes, iters, 0.0, 0,
                                        ptr(hist_d), C.byref(res), stream), 'psb_pcg_solve')
    else:
        D = pdist.DistCSR(comm, indptr, cols, data, lo, hi, n)
        del cols
        wbytes = int(lib.psb_dist_pcg_workspace_bytes(n_loc, D.n_halo))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        mode = 'nvlink-p2p' if D.p2p else 'nccl'

        def step():
            nat.check(lib.psb_dist_pcg_solve(D.handle, ptr(b_d), ptr(x_d), ptr(work), wbytes, iters, 0.0, 0,
                                             ptr(hist_d), C.byref(res), stream), 'psb_dist_pcg_solve')
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    assert res.n_hist == iters
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    nnz_t = torch.tensor([nnz_loc], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nnz_t)
    dev_ms, nnz = float(ms.item()), int(nnz_t.item())
    final = float(hist_d[-1].item())
    if world > 1:
        D.close()
    peak, peak_src = peaks()
    iter_bytes = 12 * nnz + 4 * (n + world) + 88 * n
    iter_ms = dev_ms / (steps * iters)
    gbs = iter_bytes / (iter_ms * 1e-3) / 1e9
    return {'workload': c4_name(m), 'n': n, 'nnz': nnz, 'n_gpus': world, 'steps': steps,
            'value': steps * iters / (dev_ms * 1e-3), 'unit': UNIT, 'ms_per_iteration': iter_ms,
            'collectives': mode, 'final_residual': final,
            'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak * world, 'unit': 'GB/s',
                         'frac': gbs / (peak * world), 'bytes_per_iteration': iter_bytes,
                         'peak_source': peak_src + (' x %d GPUs' % world if world > 1 else '')}}


def run_single_gpu(args):
    import ctypes as C
    import torch
    from pysolvers_b200.csrc.build import build_native
    build_native()
    from pysolvers_b200 import CommonSolverArgs, _native as nat
    from pysolvers_b200.Linear import PCG
    from pysolvers_b200.device import DeviceCSR, to_device, ptr, current_stream_ptr

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(0)
    lib = nat.lib()
    peak_gbs, peak_src = peaks()
    parity = None
    if os.environ.get('PSB_BENCH_SKIP_PARITY', '0') != '1':
        parity = parity_checks(None, 0, 1)

    A, b = build_problem(M_GRID)
    n, nnz = A.shape[0], A.nnz
    dA = DeviceCSR(A)
    dA_info = dA.info()
    b_d = to_device(b)
    x_d = torch.empty(n, dtype=torch.float64, device='cuda')
    wbytes = int(lib.psb_pcg_workspace_bytes(n, 0))
    work = torch.empty(wbytes, dtype=torch.uint8, device='cuda')
    hist_d = torch.empty(ITERS_PER_STEP, dtype=torch.float64, device='cuda')
    res = nat.SolveResult()
    stream = current_stream_ptr()

    def device_step():
        nat.check(lib.psb_pcg_solve(dA.handle, None, ptr(b_d), ptr(x_d), ptr(work), wbytes,
                                    ITERS_PER_STEP, 0.0, 0, ptr(hist_d), C.byref(res), stream))
        assert res.n_hist == ITERS_PER_STEP, res.n_hist

    # ---- value: device-resident -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        device_step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    torch.cuda.synchronize()
    launches = nat.launch_count() - l0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.finish()
    value = args.steps * ITERS_PER_STEP / (dev_ms * 1e-3)
    hist_last = float(hist_d[-1].item())

    # ---- roofline of the dominant kernel: SpMV + p.Ap ----------------------------
    p_d = torch.ones(n, dtype=torch.float64, device='cuda')
    ap_d = torch.empty(n, dtype=torch.float64, device='cuda')
    dot_d = torch.zeros(1, dtype=torch.float64, device='cuda')
    reps = 50
    for _ in range(5):
        nat.check(lib.psb_spmv_dot(dA.handle, ptr(p_d), ptr(ap_d), ptr(dot_d), stream))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        nat.check(lib.psb_spmv_dot(dA.handle, ptr(p_d), ptr(ap_d), ptr(dot_d), stream))
    e1.record()
    torch.cuda.synchronize()
    spmv_ms = e0.elapsed_time(e1) / reps
    spmv_bytes = 12 * nnz + 4 * (n + 1) + 16 * n
    spmv_gbs = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    iter_bytes = 12 * nnz + 4 * (n + 1) + 88 * n
    iter_ms = dev_ms / (args.steps * ITERS_PER_STEP)
    iter_gbs = iter_bytes / (iter_ms * 1e-3) / 1e9
    # the dominant kernel: the persistent PCG kernel, ONE launch per step (= per solve); its
    # algorithmic bytes = the init pass (read b; write r, x, p_-1: 32 n) + ITERS_PER_STEP iterations
    launches_per_step = launches / float(args.steps)
    mega = launches_per_step < 10
    step_bytes = 32 * n + ITERS_PER_STEP * iter_bytes
    step_ms = dev_ms / args.steps
    step_gbs = step_bytes / (step_ms * 1e-3) / 1e9
    # what the kernel really moves: x += alpha p is deferred into the next SpMV pass, so an iteration
    # touches 72 n (not SURVEY's 88 n) vector bytes; + the final x flush (24 n)
    moved_bytes = 32 * n + 24 * n + ITERS_PER_STEP * (12 * nnz + 4 * (n + 1) + 72 * n)
    moved_gbs = moved_bytes / (step_ms * 1e-3) / 1e9
    del p_d, ap_d

    # ---- e2e: public API with host operands ---------------------------------------
    solver = PCG(CommonSolverArgs(maxiter=ITERS_PER_STEP, tau=0.0, failOnMaxiter=False,
                                  showIters=False, showFinal=False)).makeSolver()

    def api_step():
        with contextlib.redirect_stdout(io.StringIO()):
            st = solver.solve(A, b)
        assert st.success() and st.iters() == ITERS_PER_STEP
        return st
    del dA, x_d, work
    torch.cuda.empty_cache()
    for _ in range(2):
        st = api_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = api_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = args.steps * ITERS_PER_STEP / e2e_s
    h2d = A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + b.nbytes
    d2h = 8 * n + 8 * ITERS_PER_STEP
    assert abs(st.resid() - hist_last) <= 1e-12 * abs(hist_last), (st.resid(), hist_last)

    # ---- CPU baseline: oracle port on a bounded sample ---------------------------------
    cpu_iters = int(os.environ.get('PSB_CPU_ITERS', '24'))
    t0 = time.perf_counter()
    cpu_rate, cpu_rate_1, cores = cpu_rates_both(A, b, cpu_iters)
    cpu_s = time.perf_counter() - t0

    # ---- configs[2] as named: PCG + incomplete Cholesky (RightIC defaults), at the largest size
    # whose SuperLU setup fits a bench run (m = 1024; at m = 4096 spilu alone takes ~1.7 h on the host)
    ic_line = None
    if os.environ.get('PSB_BENCH_SKIP_IC', '0') != '1':
        try:
            ic_line = ic_pcg_leg(int(os.environ.get('PSB_BENCH_IC_M', '1024')))
        except Exception as exc:                       # never lose the headline line over the side leg
            ic_line = {'error': repr(exc)[:200]}

    # ---- configs[3]: 512^3 3-D Laplacian, same step, on this one GPU --------------------------
    c4 = None
    if os.environ.get('PSB_BENCH_SKIP_C4', '0') != '1':
        del A, b, b_d, hist_d, st
        torch.cuda.empty_cache()
        try:
            c4 = c4_leg(None, 0, 1, min(args.steps, 5), 3)
        except Exception as exc:
            c4 = {'error': repr(exc)[:200]}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': bench_config(M_GRID, 1),
        'parallelism': 'single GPU', 'spmv_kernel': dA_info,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * e2e_s / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': ({'bound': 'hbm',
                      'kernel': 'pcg_mega_kernel (persistent: the whole %d-iteration solve in one launch; '
                                'launch time = CUDA-event time of the step, which also covers the '
                                'state copy-back)' % ITERS_PER_STEP,
                      'achieved': step_gbs, 'peak': peak_gbs, 'unit': 'GB/s', 'frac': step_gbs / peak_gbs,
                      'traffic': ncu_traffic('pcg_mega_kernel'),
                      'bytes_per_launch': step_bytes, 'ms_per_launch': step_ms, 'peak_source': peak_src,
                      'note': 'achieved = ALGORITHMIC bytes (SURVEY 8d: 12 nnz + 4(n+1) + 88 n per iteration) / time, so '
                              'frac can exceed 1: the kernel moves fewer bytes than that count (72 n of vectors per '
                              'iteration) -- see bytes_moved_per_launch / frac_moved for the traffic actually generated',
                      'bytes_moved_per_launch': moved_bytes, 'moved_GBps': moved_gbs, 'frac_moved': moved_gbs / peak_gbs}
                     if mega else
                     {'bound': 'hbm', 'kernel': 'spmv_bulk_kernel<EPI_DOT> (A p fused with p.Ap)',
                      'achieved': spmv_gbs, 'peak': peak_gbs, 'unit': 'GB/s',
                      'frac': spmv_gbs / peak_gbs, 'traffic': ncu_traffic('spmv_bulk_kernel'),
                      'bytes_per_launch': spmv_bytes, 'ms_per_launch': spmv_ms,
                      'peak_source': peak_src}),
        'spmv_roofline': {'bound': 'hbm', 'kernel': 'spmv_bulk_kernel<EPI_DOT> timed alone (50 launches)',
                          'achieved': spmv_gbs, 'peak': peak_gbs, 'unit': 'GB/s', 'frac': spmv_gbs / peak_gbs,
                          'traffic': ncu_traffic('spmv_bulk_kernel'), 'bytes_per_launch': spmv_bytes,
                          'ms_per_launch': spmv_ms},
        'iter_roofline': {'bound': 'hbm', 'achieved': iter_gbs, 'peak': peak_gbs, 'unit': 'GB/s',
                          'frac': iter_gbs / peak_gbs, 'bytes_per_iteration': iter_bytes,
                          'ms_per_iteration': iter_ms},
        'cpu_baseline': {'value': cpu_rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'value_1_blas_thread': cpu_rate_1,
                         'sample': '%d PCG iterations of the same %d-row system, once with OpenBLAS on %d '
                                   'thread(s) (value) and once on 1 (%.1f s together); scipy csr_matvec is '
                                   'single-threaded; os.cpu_count()=%d'
                                   % (cpu_iters, n, cores, cpu_s, os.cpu_count())},
        'final_residual': hist_last,
    }
    if parity is not None:
        line['parity'] = parity
    if c4 is not None:
        line['c4'] = c4
    if ic_line is not None:
        line['ic_pcg'] = ic_line
    print(json.dumps(line))
    return 0


def run_multi_gpu(args):
    """bench.py --gpus N under torchrun (N > 1): strong scaling of the metric workload, one rank
    per GPU; the collectives are fused into the persistent kernel over NVLink peer memory."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from pysolvers_b200.csrc.build import build_native
    from pysolvers_b200 import CommonSolverArgs, _native as nat
    from pysolvers_b200 import dist as pdist
    from pysolvers_b200.device import current_stream_ptr, ptr
    from pysolvers_b200.problems import device_fd_laplacian
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    local_rank = int(os.environ.get('LOCAL_RANK', str(rank)))
    torch.cuda.set_device(local_rank)
    build_native()
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = nat.lib()
    comm = pdist.Comm()
    dev = torch.device('cuda', local_rank)

    # ---- parity of the path about to be timed (small systems, vs scipy / the oracle) -------------
    parity = None
    if os.environ.get('PSB_BENCH_SKIP_PARITY', '0') != '1':
        parity = parity_checks(comm, rank, world)

    m = M_GRID
    n = m * m
    starts = pdist.row_starts(n, world)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    indptr, cols, data = device_fd_laplacian(2, 0.0, 1.0, m, negate=True, row_lo=lo, row_hi=hi, raw=True)
    nnz_loc = int(data.numel())
    D = pdist.DistCSR(comm, indptr, cols, data, lo, hi, n)
    n_loc = hi - lo
    b_d = torch.ones(n_loc, dtype=torch.float64, device=dev)
    x_d = torch.empty(n_loc, dtype=torch.float64, device=dev)
    iters = ITERS_PER_STEP
    wbytes = int(lib.psb_dist_pcg_workspace_bytes(n_loc, D.n_halo))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    hist_d = torch.empty(iters, dtype=torch.float64, device=dev)
    res = nat.SolveResult()
    stream = current_stream_ptr()

    def step():
        nat.check(lib.psb_dist_pcg_solve(D.handle, ptr(b_d), ptr(x_d), ptr(work), wbytes, iters, 0.0, 0,
                                         ptr(hist_d), C.byref(res), stream), 'psb_dist_pcg_solve')
        assert res.n_hist == iters

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = nat.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dev_ms = float(ms.item())
    clocks = sampler.finish() if sampler else None
    value = args.steps * iters / (dev_ms * 1e-3)
    mode = 'nvlink-p2p' if D.p2p else 'nccl'
    tile_info = D.A.info()

    # e2e: public API with host operands on every rank (upload block + b, download x).  The
    # partition plan of the structure is cached on the communicator (dist.DistCSR): each solve
    # uploads indptr / indices / values, verifies the structure on the device and reuses the plan.
    skip_e2e = os.environ.get('PSB_BENCH_SKIP_E2E', '0') == '1'
    e2e_value, h2d_total, final_resid = None, 0, float(hist_d[-1].item())
    nnz_t = torch.tensor([nnz_loc], dtype=torch.float64, device=dev)
    dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())
    if not skip_e2e:
        pin = lambda t: t.cpu().pin_memory().numpy()
        ip_h, dt_h, cols_h = pin(indptr), pin(data), pin(cols)
        b_h = torch.ones(n_loc, dtype=torch.float64).pin_memory().numpy()
        solver = pdist.DistributedPCG(CommonSolverArgs(maxiter=iters, tau=0.0, failOnMaxiter=False,
                                                       showIters=False, showFinal=False))
        D.close()
        del D, indptr, cols, data
        torch.cuda.empty_cache()

        def api_step():
            Dm = pdist.DistCSR(comm, ip_h, cols_h, dt_h, lo, hi, n)
            with contextlib.redirect_stdout(io.StringIO()):
                st = solver.solve(Dm, b_h)
            Dm.close()
            assert st.success() and st.iters() == iters
            return st
        for _ in range(2):                  # as in the single-GPU leg: the first two solves touch
            api_step()                      # freshly allocated device memory (measured 76 vs 47 ms)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = api_step()
        torch.cuda.synchronize()
        dist.barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = args.steps * iters / float(e2e_s.item())
        h2d = torch.tensor([ip_h.nbytes + cols_h.nbytes + dt_h.nbytes + b_h.nbytes], dtype=torch.float64, device=dev)
        dist.all_reduce(h2d)
        h2d_total = int(h2d.item())
        final_resid = float(st.resid())
        del ip_h, dt_h, cols_h
    else:
        D.close()
        del D, indptr, cols, data
    del x_d, work, b_d
    torch.cuda.empty_cache()

    # ---- configs[3]: 512^3 3-D Laplacian on the same ranks ---------------------------------------
    c4 = None
    if os.environ.get('PSB_BENCH_SKIP_C4', '0') != '1':
        c4 = c4_leg(comm, rank, world, min(args.steps, 5), 3)

    if rank == 0:
        peak, peak_src = peaks()
        iter_bytes = 12 * nnz + 4 * (n + world) + 88 * n
        iter_ms = dev_ms / (args.steps * iters)
        gbs = iter_bytes / (iter_ms * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': bench_config(m, world),
            'parallelism': 'row partition over %d GPUs, collectives: %s' % (world, mode),
            'collectives': mode, 'spmv_kernel': tile_info,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_total,
                    'd2h_bytes_per_step': int(8 * n + 8 * iters * world)},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'pcg_mega_kernel: whole PCG iteration (aggregate over GPUs)',
                         'achieved': gbs, 'peak': peak * world, 'unit': 'GB/s',
                         'frac': gbs / (peak * world), 'traffic': None,
                         'bytes_per_iteration': iter_bytes, 'ms_per_iteration': iter_ms,
                         'peak_source': peak_src + ' x %d GPUs' % world},
            'cpu_baseline': None,
            'final_residual': final_resid,
        }
        if parity is not None:
            line['parity'] = parity
        if c4 is not None:
            line['c4'] = c4
        print(json.dumps(line))
    comm.close()
    dist.destroy_process_group()
    return 0


def ic_pcg_leg(m):
    """IC-preconditioned PCG to tau = 1e-8 through the public API on the m x m Laplacian; the
    preconditioner application is timed alone on the device and with scipy on the host."""
    import torch
    from oracle import precond
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import PCG, RightIC
    from pysolvers_b200.device import to_device
    from pysolvers_b200.problems import fd_laplacian_2d
    A = -fd_laplacian_2d(0.0, 1.0, m)
    n = A.shape[0]
    b = np.ones(n)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pre = RightIC().form(A)
    setup_s = time.perf_counter() - t0
    dev = pre.device_prec()
    v, z = to_device(b), torch.empty(n, dtype=torch.float64, device='cuda')
    for _ in range(2):
        dev.apply(v, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dev.apply(v, z)
    e1.record()
    torch.cuda.synchronize()
    apply_ms = e0.elapsed_time(e1) / 5
    iL, iLt = pre._dL.info(), pre._dLt.info()
    t0 = time.perf_counter()
    ref = precond.ic_apply(pre._L, pre._Lt, b)
    cpu_apply_s = time.perf_counter() - t0
    err = float(np.linalg.norm(z.cpu().numpy() - ref) / np.linalg.norm(ref))
    s = PCG(CommonSolverArgs(maxiter=2000, tau=1e-8, showIters=False, showFinal=False), precond=RightIC()).makeSolver()
    s.precond = pre
    s.freezePrec()                                    # reuse the factor formed above (PCGSolver.py:92-94)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        st = s.solve(A, b)
    torch.cuda.synchronize()
    solve_s = time.perf_counter() - t0
    levels = iL['levels'] + iLt['levels']
    return {'workload': 'IC-PCG (RightIC defaults), 2-D 5-point Laplacian m=%d (n=%d), tau=1e-8' % (m, n),
            'iterations': int(st.iters()), 'success': bool(st.success()),
            'iter_per_s_e2e': st.iters() / solve_s, 'solve_s': solve_s,
            'ic_setup_host_s': setup_s, 'nnz_L': int(pre._L.nnz), 'levels_L_plus_Lt': int(levels),
            'ic_apply_ms': apply_ms, 'us_per_level': 1e3 * apply_ms / levels,
            'ic_apply_kernel': pre._dL.info2()['kernel'],
            'ic_apply_algorithmic_GBps': (12 * (iL['nnz_packed'] + iLt['nnz_packed']) + 48 * n) / (apply_ms * 1e-3) / 1e9,
            'cpu_ic_apply_s': cpu_apply_s, 'apply_rel_err_vs_scipy': err}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.gpus == 1 and int(os.environ.get('WORLD_SIZE', '1')) == 1:
        return run_single_gpu(args)
    return run_multi_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
