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
    A, b = build_problem(M_GRID, pinned=False)
    iters = int(os.environ.get('PSB_REF_ITERS_PER_STEP', '4'))   # bounded sample per step
    for _ in range(args.warmup):
        cpu_oracle_rate(A, b, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_rate(A, b, iters)
    dt = time.perf_counter() - t0
    value = args.steps * iters / dt
    cores = blas_threads()
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(M_GRID), 'sample': '%d iterations per step' % iters},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d steps x %d PCG iterations of the full %d-row system; '
                                   'scipy csr_matvec is single-threaded, OpenBLAS ddot uses %d thread(s); '
                                   'os.cpu_count()=%d' % (args.steps, iters, A.shape[0], cores, os.cpu_count())},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


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
    cpu_rate, cpu_s = cpu_oracle_rate(A, b, cpu_iters)
    cores = blas_threads()

    # ---- configs[2] as named: PCG + incomplete Cholesky (RightIC defaults), at the largest size
    # whose SuperLU setup fits a bench run (m = 1024; at m = 4096 spilu alone takes ~1.7 h on the host)
    ic_line = None
    if os.environ.get('PSB_BENCH_SKIP_IC', '0') != '1':
        try:
            ic_line = ic_pcg_leg(int(os.environ.get('PSB_BENCH_IC_M', '1024')))
        except Exception as exc:                       # never lose the headline line over the side leg
            ic_line = {'error': repr(exc)[:200]}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(M_GRID), 'n': n, 'nnz': nnz,
                   'iters_per_step': ITERS_PER_STEP, 'parallelism': 'single GPU',
                   'l2': 'inputs larger than L2: %.2f GB touched per iteration vs 126 MB L2'
                         % (iter_bytes / 1e9),
                   'spmv_kernel': dA_info},
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
                      'bytes_per_launch': step_bytes, 'ms_per_launch': step_ms, 'peak_source': peak_src}
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
                         'sample': '%d PCG iterations of the same %d-row system (%.1f s); scipy '
                                   'csr_matvec single-threaded, OpenBLAS %d thread(s), os.cpu_count()=%d'
                                   % (cpu_iters, n, cpu_s, cores, os.cpu_count())},
        'final_residual': hist_last,
    }
    if ic_line is not None:
        line['ic_pcg'] = ic_line
    print(json.dumps(line))
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
    if (args.gpus == 1 and int(os.environ.get('WORLD_SIZE', '1')) == 1
            and os.environ.get('PSB_BENCH_WORKLOAD', 'c3') == 'c3'):
        return run_single_gpu(args)
    from pysolvers_b200.dist import bench_multi_gpu
    return bench_multi_gpu(args, sys.modules[__name__])


if __name__ == '__main__':
    sys.exit(main())
