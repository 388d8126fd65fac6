#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: Circulant(N = 2^20).forward on 1024 complex64 columns per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (``Circulant.forward``) over one batch of synthetic columns.  The column batch
shards across GPUs with no data-path collective (weak scaling: every rank transforms its own 1024 columns); under
torchrun each rank times its own steps with CUDA events between barriers and rank 0 reports the max over ranks.

JSON keys beyond the base contract:
  roofline     algorithmic HBM bytes of the operator (2 * 8 B * N per column) / CUDA-event time, against the
               measured copy bandwidth of MEASURED_PEAKS.json; "kernels" gives each pass's share of a step
  e2e          the same metric through the host-buffer API (Matrix.apply_host): pinned host input -> H2D -> transform ->
               D2H, all inside the timed region, every step, on the full column batch; per-GPU copy rates and the copy-only
               ceiling of the same buffers measured in the same run
  sustained    the headline step repeated back to back for >= 2 s (power-capped steady state) next to the K-step figure
  output_check two columns of the timed batch against the reference evaluated in double precision (outside timed regions)
  cpu_baseline the reference's own CPU implementation (oracle/_ref, built from /root/reference) on this box's host
               cores, on a bounded column sample
  extras       other operators of the path at their BASELINE shapes (not the headline, same timing rules)
`--impl reference` times the reference's CPU path (all host cores, column shards in worker processes).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ORDER = 20
N = 1 << N_ORDER
COLS = 1024
METRIC = "circulant_forward_columns_per_s"
UNIT = "columns/s"


def source_hash():
    """sha256 over the kernel sources (csrc + the C-ABI header): ties profile-derived numbers (profiles/r2_traffic.json) to
    the code that was timed - the library itself is rebuilt on every box, its bytes are not comparable."""
    import hashlib
    h = hashlib.sha256()
    src = os.path.join(ROOT, 'fastmat_b200', 'csrc')
    for f in sorted(os.listdir(src)) + ['../../include/fastmat_b200.h']:
        path = os.path.normpath(os.path.join(src, f))
        if os.path.isfile(path):
            h.update(f.encode())
            with open(path, 'rb') as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for k in ('HwSlowdown', 'HwThermalSlowdown', 'SwThermalSlowdown', 'SwPowerCap', 'HwPowerBrakeSlowdown'):
            v = getattr(nv, 'nvmlClocksEventReason' + k, None) or getattr(nv, 'nvmlClocksThrottleReason' + k, None)
            if v is not None:
                names[v] = k
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons', None)
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                if get_reasons is not None:
                    r = get_reasons(self.h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        to_snake = {'HwSlowdown': 'hw_slowdown', 'HwThermalSlowdown': 'hw_thermal_slowdown',
                    'SwThermalSlowdown': 'sw_thermal_slowdown', 'SwPowerCap': 'sw_power_cap',
                    'HwPowerBrakeSlowdown': 'hw_power_brake_slowdown'}
        return {'sm_mhz': (s[len(s) // 2] if s else None), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(to_snake.get(r, r) for r in self.reasons), 'samples': len(s)}


# ------------------------------------------------------------------------------------------- CPU reference
def _load_reference():
    """The real reference (oracle/_ref) if it imports, else the numpy port (oracle/fastmat_oracle.py)."""
    ref_dir = os.path.join(ROOT, 'oracle', '_ref')
    try:
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        import fastmat
        return 'reference', fastmat
    except Exception:
        from oracle import fastmat_oracle
        return 'port', fastmat_oracle


def _cpu_worker(args):
    kind_hint, n, cols, seed, reps = args
    import numpy as np
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    kind, mod = _load_reference()
    rng = np.random.default_rng(4321)
    c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    rng = np.random.default_rng(seed)
    x = np.asfortranarray((rng.standard_normal((n, cols)) + 1j * rng.standard_normal((n, cols))).astype(np.complex64))
    if kind == 'reference':
        C = mod.Circulant(c)
        f = lambda: C.forward(x)            # noqa: E731
    else:
        f = lambda: mod.circulant_forward(c, x, double=True)   # noqa: E731
    f()                                      # warm-up (plan construction, page faults)
    best = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        best.append(time.perf_counter() - t0)
    return kind, best


def cpu_baseline_single(cols=16, reps=2):
    """1 core, the way the reference actually runs (single-threaded pocketfft + Cython loops)."""
    kind, times = _cpu_worker(('', N, cols, 1234, reps))
    t = min(times)
    return {'value': cols / t, 'unit': UNIT, 'cores': 1, 'kind': kind,
            'sample': 'Circulant(2^20).forward on %d complex64 columns (F-order), best of %d, 1 process' % (cols, reps)}


def check_against_reference(c, x_cols, y_cols):
    """Part of the cpu_baseline leg (the checker, outside every timed region): the reference's Circulant.forward evaluated in
    double precision (generator and input cast to complex128, SURVEY 8c rule 1) on a few columns of the timed batch,
    compared with the GPU output of the same columns.  Returns the error normalised as the tests do:
    max|y - y_ref| / (||x||_2 * log2 N) per column, worst column."""
    import numpy as np
    kind, mod = _load_reference()
    c128 = c.astype(np.complex128)
    x128 = np.asfortranarray(x_cols.astype(np.complex128))
    if kind == 'reference':
        ref = mod.Circulant(c128).forward(x128)
    else:
        ref = mod.circulant_forward(c128, x128, double=True)
    err = np.abs(y_cols.astype(np.complex128) - ref).max(axis=0)
    norm = np.sqrt((np.abs(x128) ** 2).sum(axis=0)) * N_ORDER
    return {'columns_checked': int(x_cols.shape[1]), 'max_err_over_norm_x_log2n': float((err / norm).max()),
            'tolerance': 1e-5, 'reference': kind + ' in complex128', 'ok': bool((err / norm).max() <= 1e-5)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation on all host cores (one worker process per core, disjoint
    column shards), same metric / config; each step is a bounded sample of the workload."""
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    cols_per_worker = 4
    ctx = mp.get_context('spawn')
    steps = max(1, args.steps)
    warm = max(0, args.warmup)
    t_steps = []
    kind = 'port'
    with ctx.Pool(workers) as pool:
        jobs = [('', N, cols_per_worker, 1000 + w, 1) for w in range(workers)]
        for s in range(min(warm, 1) + min(steps, 3)):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
            wall = time.perf_counter() - t0
            kind = res[0][0]
            # throughput of the step = columns of all workers / slowest worker's transform time
            slow = max(min(r[1]) for r in res)
            if s >= min(warm, 1):
                t_steps.append(slow)
            del wall
    t = sum(t_steps) / len(t_steps)
    value = workers * cols_per_worker / t
    sample = ('Circulant(2^20).forward, %d worker processes x %d complex64 columns per step (F-order), %d timed steps'
              % (workers, cols_per_worker, len(t_steps)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(t_steps),
        'warmup': min(warm, 1), 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'complex64 in / complex128 internally (reference Product.pyx:215)', 'data': 'synthetic',
        'config': {'workload': 'Circulant(N=2^20).forward, complex64, %d columns per GPU, column-major (fastmat layout)' % COLS,
                   'n': N, 'columns_per_gpu': COLS, 'sample': sample,
                   'note': 'reference arm: fastmat CPU implementation (oracle/_ref) on the host cores, bounded column sample per step'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': workers, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------- config 5 (solvers)
def run_config5(fm, fpar, dev, rank, world, distributed, barrier, max_over_ranks, cols_per_gpu):
    """fastmat.algorithms ISTA (100 steps) and OMP (k = 32) on the compressed-sensing operator of BASELINE config 5, one
    shard of right-hand sides per GPU through parallel.solve_sharded; reports columns solved per second (whole job) and
    checks the recovery (OMP: exact support)."""
    import numpy as np
    import torch
    n, m, k = 1 << 18, 1 << 16, 32
    rng = np.random.default_rng(2026)                                    # the operator is the same on every rank
    rows = np.sort(rng.choice(n, m, replace=False))
    d = np.exp(2j * np.pi * rng.random(n)).astype(np.complex64)
    A = fm.Product(fm.Partial(fm.Fourier(n), rows=rows), fm.Diag(d))
    L = cols_per_gpu
    rl = np.random.default_rng(7000 + rank)                              # every rank has its own right-hand sides
    x = np.zeros((L, n), dtype=np.complex64)
    for c in range(L):
        idx = rl.choice(n, k, replace=False)
        x[c, idx] = (2 + rl.random(k)) * np.exp(2j * np.pi * rl.random(k))
    xd = torch.from_numpy(x).to(dev).t()                                 # column-major (n, L)
    b = A.forward(xd)
    out = {'operator': 'Product(Partial(Fourier(2^18), 2^16 sorted random rows), Diag(unit modulus)), complex64',
           'columns_per_gpu': L, 'global_columns': L * world, 'sparsity': k,
           'sharding': 'parallel.solve_sharded: one shard of right-hand sides per GPU, no collective in the iteration'}

    def wall(fn):
        barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        return r, dt

    ista = fm.algorithms.ISTA(A, numLambda=50.0, numMaxSteps=100)
    fpar.solve_sharded(fm.algorithms.ISTA(A, numLambda=50.0, numMaxSteps=2), b) if distributed else ista.process(b, numMaxSteps=2)
    ista.numMaxSteps = 100
    res, dt = wall(lambda: (fpar.solve_sharded(ista, b) if distributed else ista.process(b)))
    top = torch.topk(res.abs(), k, dim=0).indices
    hit = float(((xd != 0).gather(0, top)).double().mean().item())
    out['ista_100_steps'] = {'seconds': dt, 'columns_per_s': L * world / dt, 'operator_applies_per_s': 200 * L * world / dt,
                             'top_k_on_true_support': hit}
    omp = fm.algorithms.OMP(A, numMaxSteps=2)
    omp.process(b)                                                       # warm-up: colNormalized, double-precision plans, solver handles
    omp.numMaxSteps = k
    res, dt = wall(lambda: (fpar.solve_sharded(omp, b, share_step_size=False) if distributed else omp.process(b)))
    out['omp_k32'] = {'seconds': dt, 'columns_per_s': L * world / dt,
                      'support_exact': bool(torch.equal(res != 0, xd != 0)),
                      'max_abs_error': float((res - xd).abs().max().item())}
    if distributed:
        full, dt = wall(lambda: fpar.gather_columns(res, L * world))
        out['gather_columns_all_gather'] = {'seconds': dt, 'bytes_per_rank': int(res.numel() * res.element_size()),
                                            'gbs_per_rank_received': res.numel() * res.element_size() * (world - 1) / dt / 1e9,
                                            'shape': list(full.shape)}
    return out


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import fastmat_b200 as fm

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'       # NCCL prints its version banner on stdout; stdout carries the JSON line only
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not distributed:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cols = args.cols
    steps, warm = max(1, args.steps), max(3, args.warmup)
    from fastmat_b200 import parallel as fpar
    cores = fpar.bind_to_gpu_numa(local_rank) if distributed else None      # before any pinned allocation
    rng = np.random.default_rng(4321)
    c = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    C = fm.Circulant(c)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def crandn(rows, m):
        """column-major (rows, m) complex64 standard-normal tensor (fastmat's native layout)."""
        t = torch.empty((m, rows), dtype=torch.complex64, device=dev)
        tr = torch.view_as_real(t)
        tr.normal_(generator=g)
        return t.t()

    x = crandn(N, cols)

    def timed(fn, k, w, sustain_s=0.0):
        """ms per step over k steps (max over ranks); with sustain_s > 0 additionally the mean over a back-to-back run of at
        least that many seconds (the power-capped steady state)."""
        for _ in range(w):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / k          # ms per step, max over ranks
        if sustain_s <= 0:
            return ms, None
        n = max(k, int(sustain_s * 1e3 / ms) + 1)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return ms, max_over_ranks(e0.elapsed_time(e1)) / n

    # ---- headline: device-resident inputs
    sampler = ClockSampler(local_rank)
    for _ in range(warm):
        y = C.forward(x)
    barrier()
    launches0 = fm.launch_count()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y = C.forward(x)
    e1.record()
    barrier()
    launches = fm.launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    # the same loop back to back for >= args.sustain seconds: what the operator holds under the power cap
    n_sus = max(steps, int(args.sustain * 1e3 / ms) + 1) if args.sustain > 0 else 0
    ms_sus = None
    if n_sus:
        e0.record()
        for _ in range(n_sus):
            y = C.forward(x)
        e1.record()
        barrier()
        ms_sus = max_over_ranks(e0.elapsed_time(e1)) / n_sus
    clocks = sampler.stop()
    value = world * cols / (ms * 1e-3)
    peak, peak_src = measured_peak()
    alg_bytes = 2.0 * 8.0 * N * cols                             # read x + write y, per GPU and step
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    info = C._plan.info

    # per-kernel share of a step and DRAM traffic come from the ncu launch list of this command committed under profiles/;
    # they are quoted only if that list was captured on the kernel sources being timed now (source hash), else null + why
    traffic, kernels, dominant, prof_note, traffic_conc, conc_note = None, None, None, None, None, None
    shash = source_hash()
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')) as f:
            tj = json.load(f)
        if tj.get('source_hash') != shash:
            prof_note = 'profiles/r2_traffic.json was captured on other kernel sources (%s, timed now: %s): not quoted' % (tj.get('source_hash'), shash)
        else:
            if cols == COLS:
                traffic = tj['dram_bytes_per_step']
                traffic_conc = tj.get('dram_bytes_per_step_concurrent')
                conc_note = tj.get('concurrent_note')
            kernels = {k: {'share_of_step_time': v['share_of_step_time'], 'avg_us_under_ncu': v['avg_us']}
                       for k, v in tj['kernels'].items() if v['share_of_step_time'] > 0.01}
            dominant = tj.get('dominant_kernel')
            prof_note = 'from profiles/r2_traffic.json (ncu launch list of this command, same kernel sources: %s)' % shash
    except Exception as e:
        prof_note = 'profiles/r2_traffic.json not readable (%s)' % (e.__class__.__name__, )

    # columns of the timed batch kept for the check against the reference (cpu_baseline leg, rank 0)
    chk_cols = 2
    x_chk = x[:, :chk_cols].cpu().numpy() if rank == 0 else None
    y_chk = y[:, :chk_cols].cpu().numpy() if rank == 0 else None
    y_dtype = y.dtype
    del y

    # ---- e2e: host buffers through Matrix.apply_host (H2D + transform + D2H inside the timed region, every step)
    e2e = None
    if not args.no_e2e:
        e2e_cols = min(cols, args.e2e_cols)
        avail_gib = 0.0
        try:
            with open('/proc/meminfo') as f:
                for ln in f:
                    if ln.startswith('MemAvailable'):
                        avail_gib = float(ln.split()[1]) / 2 ** 20
        except Exception:
            pass
        need_gib = 2 * 8.0 * N * e2e_cols / 2 ** 30 * world
        while avail_gib and need_gib > 0.5 * avail_gib and e2e_cols > 64:      # pinned staging must fit the host comfortably
            e2e_cols //= 2
            need_gib /= 2
        xh = torch.empty((e2e_cols, N), dtype=torch.complex64, pin_memory=True).t()
        xh.copy_(x[:, :e2e_cols])
        yh = torch.empty((e2e_cols, N), dtype=y_dtype, pin_memory=True).t()
        k_e2e = max(1, min(steps, args.e2e_steps))
        st = {}
        C.apply_host(xh, out=yh)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            C.apply_host(xh, out=yh, stats=st)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        # copy-only ceiling of the same buffers (nothing computed): what the host path could reach if the transform were free
        barrier()
        ct, cb_in, cb_out = fpar.host_copy_ceiling(xh, yh, dev)
        ct = max_over_ranks(ct)
        e2e = {'value': world * e2e_cols * k_e2e / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(st['h2d_bytes']),
               'd2h_bytes_per_step': int(st['d2h_bytes']), 'columns_per_step_per_gpu': e2e_cols, 'steps': k_e2e,
               'h2d_gbs_per_gpu': st['h2d_bytes'] * k_e2e / dt / 1e9, 'd2h_gbs_per_gpu': st['d2h_bytes'] * k_e2e / dt / 1e9,
               'copy_only_ceiling': {'value': world * e2e_cols / ct, 'unit': UNIT, 'h2d_gbs_per_gpu': cb_in / ct / 1e9,
                                     'd2h_gbs_per_gpu': cb_out / ct / 1e9,
                                     'note': 'same pinned buffers and chunking, both directions at once, no transform; max over ranks'},
               'fraction_of_copy_ceiling': (world * e2e_cols * k_e2e / dt) / (world * e2e_cols / ct),
               'host_cores_bound': (len(cores) if cores else None),
               'api': 'Circulant.apply_host(pinned CPU tensor) == Circulant.forward(host array)'}
        del xh, yh

    # ---- extras: the other operators of the path at their BASELINE shapes, 20-step and sustained figures
    extras = {}
    if rank == 0 and not args.quick and not distributed:
        k2 = max(3, steps // 2)
        sus = min(args.sustain, 1.0)

        def rec(name, fn, ncols, bytes_per_col, k=None, w=3, sustain=None):
            try:
                # every row starts from an idle GPU (the rows before it leave the board power-capped): ms_per_step is the
                # figure of the operator timed alone, sustained_ms_per_step the back-to-back steady state next to it
                torch.cuda.synchronize()
                time.sleep(args.extras_idle)
                ms_, ms_s = timed(fn, k or k2, w, sus if sustain is None else sustain)
                r = {'ms_per_step': ms_, 'columns': ncols, 'columns_per_s': ncols / (ms_ * 1e-3),
                     'hbm_gbs': bytes_per_col * ncols / (ms_ * 1e-3) / 1e9,
                     'roofline_frac': bytes_per_col * ncols / (ms_ * 1e-3) / 1e9 / peak}
                if ms_s is not None:
                    r['sustained_ms_per_step'] = ms_s
                    r['sustained_roofline_frac'] = bytes_per_col * ncols / (ms_s * 1e-3) / 1e9 / peak
                extras[name] = r
            except Exception as e:                                  # a secondary row must not take the headline down
                extras[name] = {'error': repr(e)}
                torch.cuda.synchronize()

        rec('circulant_backward_2^20_c64', lambda: C.backward(x), cols, 16.0 * N)
        F = fm.Fourier(N)
        rec('fourier_forward_2^20_c64', lambda: F.forward(x), cols, 16.0 * N)
        K = fm.Kron(fm.Fourier(1024), fm.Fourier(1024))
        rec('kron_fourier_1024x1024_c64', lambda: K.forward(x), cols, 16.0 * N)
        # torch's default row-major layout (batch contiguous): SURVEY 8d asks for both layouts
        xr = x.contiguous()
        rec('circulant_forward_2^20_c64_row_major', lambda: C.forward(xr), cols, 16.0 * N, k=2, w=1, sustain=0)
        rec('fourier_forward_2^20_c64_row_major', lambda: F.forward(xr), cols, 16.0 * N, k=2, w=1, sustain=0)
        del xr
        nt = 1 << 19
        vc = (rng.standard_normal(nt) + 1j * rng.standard_normal(nt)).astype(np.complex64)
        vr = (rng.standard_normal(nt - 1) + 1j * rng.standard_normal(nt - 1)).astype(np.complex64)
        T = fm.Toeplitz(vc, vr)
        xt = x[:nt, :].t().contiguous().t()
        rec('toeplitz_forward_2^19_c64', lambda: T.forward(xt), cols, 8.0 * 2 * nt)
        rec('toeplitz_backward_2^19_c64', lambda: T.backward(xt), cols, 8.0 * 2 * nt)
        del xt
        Fb = fm.Fourier(1000003)
        xb = x[:1000003, :].t().contiguous().t()
        rec('fourier_bluestein_1000003_c64', lambda: Fb.forward(xb), cols, 16.0 * 1000003, k=3, w=2)
        del xb
        Hd = fm.Hadamard(20)
        xf = torch.empty((4 * cols, N), dtype=torch.float32, device=dev).normal_(generator=g).t()   # (2^20, 4096) float32, column-major
        rec('hadamard_forward_o20_f32', lambda: Hd.forward(xf), xf.shape[1], 8.0 * N)
        try:            # scatter -> FWHT(order 20) -> gather; 2 x 4 B x (2^20 - 1) per column
            Ll = fm.LFSRCirculant((1 << 20) | (1 << 3) | 1, 1)
            xl = xf[:N - 1, :1024].t().contiguous().t()
            rec('lfsr_circulant_forward_o20_f32', lambda: Ll.forward(xl), xl.shape[1], 8.0 * (N - 1))
            del xl, Ll
        except Exception as e:
            extras['lfsr_circulant_forward_o20_f32'] = {'error': repr(e)}
        del xf
        # mid sizes: 2^13 ... 2^15 (pass lengths 64 / 128) run the specialised passes since round 2; the 2^a 3^b lengths
        # fastmat's planner likes (6144, 110592) run as chirp-z transforms over the next power of two
        # (specialised kernels) instead of the run-time-radix kernels (0.225 / 2.15 ms)
        for nn, mm, tag in ((1 << 13, 1024, 'fast_path'), (1 << 14, 1024, 'fast_path'), (1 << 15, 1024, 'fast_path'),
                            (6144, 1024, 'chirp_z_pow2'), (110592, 256, 'chirp_z_pow2')):
            Fs = fm.Fourier(nn)
            xs_ = crandn(nn, mm)
            rec('fourier_forward_%d_c64_%s' % (nn, tag), lambda: Fs.forward(xs_), mm, 16.0 * nn, k=10, sustain=0)
            del xs_, Fs
        # lengths that fit on chip: ONE kernel per apply (FFT -> spectrum -> FFT in shared memory for Circulant / Toeplitz,
        # zero padding folded into the loads); batch of 2^26 elements so that the operands exceed L2
        for nn in (1 << 10, 1 << 12):
            mm = (1 << 26) // nn
            xs_ = crandn(nn, mm)
            cs_ = (rng.standard_normal(nn) + 1j * rng.standard_normal(nn)).astype(np.complex64)
            Fs, Cs = fm.Fourier(nn), fm.Circulant(cs_)
            Ts = fm.Toeplitz(cs_[:nn // 2], cs_[nn // 2:nn - 1])
            rec('fourier_forward_%d_c64_single_kernel' % nn, lambda: Fs.forward(xs_), mm, 16.0 * nn, k=10, sustain=0)
            rec('circulant_forward_%d_c64_single_kernel' % nn, lambda: Cs.forward(xs_), mm, 16.0 * nn, k=10, sustain=0)
            xt_ = xs_[:nn // 2, :]
            rec('toeplitz_forward_%d_c64_single_kernel' % (nn // 2), lambda: Ts.forward(xt_), mm, 8.0 * nn, k=10, sustain=0)
            xr_ = xs_.contiguous()
            rec('fourier_forward_%d_c64_single_kernel_row_major' % nn, lambda: Fs.forward(xr_), mm, 16.0 * nn, k=10, sustain=0)
            del xs_, xt_, xr_, Fs, Cs, Ts
        x16 = crandn(1 << 16, 64).to(torch.complex128)
        F16 = fm.Fourier(1 << 16)
        rec('fourier_forward_2^16_c128_64cols', lambda: F16.forward(x16), 64, 32.0 * (1 << 16), k=20)
        del x16

    # ---- BASELINE config 5: ISTA / OMP on Product(Partial(Fourier(2^18), 2^16 rows), Diag), the 1024 right-hand sides of
    #      the 8-GPU batch sharded 128 per GPU (every rank solves its own shard: no collective inside the solvers; the
    #      assembled result costs one all_gather, timed separately).  Runs at every N.
    config5 = None
    if not args.quick or distributed:
        try:
            config5 = run_config5(fm, fpar, dev, rank, world, distributed, barrier, max_over_ranks, args.c5_cols)
        except Exception as e:
            config5 = {'error': repr(e)}
            torch.cuda.synchronize()

    cpu = None
    check = None
    if rank == 0 and not args.no_cpu and not distributed:          # reported at N = 1 only
        cpu = cpu_baseline_single(cols=args.cpu_cols, reps=2)
    if rank == 0 and not args.no_cpu:
        try:
            check = check_against_reference(c, x_chk, y_chk)
        except Exception as e:
            check = {'error': repr(e)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warm,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex64',
            'data': 'synthetic',
            'config': {'workload': 'Circulant(N=2^20).forward, complex64, %d columns per GPU, column-major (fastmat layout)' % cols,
                       'n': N, 'columns_per_gpu': cols, 'global_columns': world * cols,
                       'parallelism': 'column shards, no collective on the apply path',
                       'l2': 'inputs per step (%.1f GiB) exceed the 126 MB L2, no flush needed' % (8.0 * N * cols / 2 ** 30),
                       'inner_fft': int(info.inner_size), 'passes_per_slab': int(info.passes_fwd), 'slab_cols': int(info.slab_cols)},
            'clocks': clocks,
            'sustained': (None if ms_sus is None else
                          {'ms_per_step': ms_sus, 'steps': n_sus, 'value': world * cols / (ms_sus * 1e-3), 'unit': UNIT,
                           'roofline_frac': alg_bytes / (ms_sus * 1e-3) / 1e9 / peak,
                           'note': 'the same step repeated back to back for >= %.1f s (power-capped steady state)' % args.sustain}),
            'e2e': e2e,
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_under_concurrency': traffic_conc, 'traffic_under_concurrency_note': conc_note,
                         'peak_source': peak_src,
                         'algorithmic_bytes_per_step': alg_bytes,
                         'note': ('operator level: algorithmic bytes of one step (2 x 8 B x N per column, SURVEY 8d) / CUDA-event '
                                  'time of the step on the caller stream; one step = %d launches (3 passes per slab of %d columns, '
                                  'slabs issued round-robin on internal streams so a per-kernel event time does not exist); '
                                  'kernels / traffic: ' % (launches // steps, int(info.slab_cols))) + str(prof_note),
                         'dominant_kernel': dominant,
                         'launches_per_step': launches // steps,
                         'kernels': kernels,
                         'kernel_source_hash': shash},
            'output_check': check,
            'cpu_baseline': cpu,
            'config5_solvers': config5,
            'extras': extras,
        }
        emit(line)
    if distributed:
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cols', type=int, default=COLS)
    ap.add_argument('--e2e-cols', type=int, default=COLS, dest='e2e_cols', help='columns per GPU of the host-buffer (e2e) step')
    ap.add_argument('--e2e-steps', type=int, default=2, dest='e2e_steps')
    ap.add_argument('--sustain', type=float, default=2.0, help='seconds of back-to-back steps for the sustained figure (0: off)')
    ap.add_argument('--cpu-cols', type=int, default=16, dest='cpu_cols')
    ap.add_argument('--c5-cols', type=int, default=128, dest='c5_cols', help='right-hand sides per GPU of the config-5 solver rows')
    ap.add_argument('--extras-idle', type=float, default=1.0, dest='extras_idle', help='idle seconds before each extras row')
    ap.add_argument('--quick', action='store_true', help='headline only (no extras)')
    ap.add_argument('--no-e2e', action='store_true', dest='no_e2e')
    ap.add_argument('--no-cpu', action='store_true', dest='no_cpu')
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: whatever libraries write to file descriptor 1 (NCCL's version
    # banner, for one) goes to stderr; the line itself is written to a private duplicate of the original stdout
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
