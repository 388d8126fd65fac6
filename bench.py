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
               D2H, all inside the timed region, every step
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
    print(json.dumps(line))
    return 0


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

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / k          # ms per step, max over ranks

    # ---- headline: device-resident inputs
    sampler = ClockSampler(local_rank)
    launches0 = fm.launch_count()
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
    clocks = sampler.stop()
    launches = fm.launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    value = world * cols / (ms * 1e-3)
    peak, peak_src = measured_peak()
    alg_bytes = 2.0 * 8.0 * N * cols                             # read x + write y, per GPU and step
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    info = C._plan.info

    # per-kernel share of a step and DRAM traffic: from the committed ncu launch list of this same command
    # (profiles/r1_traffic.json <- profiles/r1_launches_bench_circulant.csv); live numbers above are CUDA events only
    traffic = None
    kernels = {}
    dominant = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as f:
            tj = json.load(f)
        if cols == COLS:
            traffic = tj['dram_bytes_per_step']
        kernels = {k: {'share_of_step_time': v['share_of_step_time'], 'avg_us_under_ncu': v['avg_us']}
                   for k, v in tj['kernels'].items() if v['share_of_step_time'] > 0.01}
        dominant = tj.get('dominant_kernel')
    except Exception:
        pass
    del y

    # ---- e2e: host buffers through Matrix.apply_host (H2D + transform + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_cols = min(cols, args.e2e_cols)
        xh = torch.empty((e2e_cols, N), dtype=torch.complex64, pin_memory=True).t()
        xh.copy_(x[:, :e2e_cols])
        yh = torch.empty((e2e_cols, N), dtype=torch.complex64, pin_memory=True).t()
        k_e2e = max(1, min(steps, 3))
        C.apply_host(xh, out=yh)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            C.apply_host(xh, out=yh)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt)
        e2e = {'value': world * e2e_cols * k_e2e / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(8 * N * e2e_cols),
               'd2h_bytes_per_step': int(8 * N * e2e_cols), 'columns_per_step_per_gpu': e2e_cols, 'steps': k_e2e,
               'api': 'Circulant.apply_host(pinned CPU tensor) == Circulant.forward(host array)'}
        del xh, yh

    # ---- extras: the other operators of the path at their BASELINE shapes
    extras = {}
    if rank == 0 and not args.quick and not distributed:
        k2 = max(3, steps // 2)

        def rec(name, ms_, ncols, bytes_per_col):
            extras[name] = {'ms_per_step': ms_, 'columns_per_s': ncols / (ms_ * 1e-3),
                            'hbm_gbs': bytes_per_col * ncols / (ms_ * 1e-3) / 1e9,
                            'roofline_frac': bytes_per_col * ncols / (ms_ * 1e-3) / 1e9 / peak}
        rec('circulant_backward_2^20_c64', timed(lambda: C.backward(x), k2, 3), cols, 16.0 * N)
        F = fm.Fourier(N)
        rec('fourier_forward_2^20_c64', timed(lambda: F.forward(x), k2, 3), cols, 16.0 * N)
        nt = 1 << 19
        vc = (rng.standard_normal(nt) + 1j * rng.standard_normal(nt)).astype(np.complex64)
        vr = (rng.standard_normal(nt - 1) + 1j * rng.standard_normal(nt - 1)).astype(np.complex64)
        T = fm.Toeplitz(vc, vr)
        xt = x[:nt, :]
        xt = xt.t().contiguous().t()
        rec('toeplitz_forward_2^19_c64', timed(lambda: T.forward(xt), k2, 3), cols, 8.0 * 2 * nt)
        rec('toeplitz_backward_2^19_c64', timed(lambda: T.backward(xt), k2, 3), cols, 8.0 * 2 * nt)
        del xt
        K = fm.Kron(fm.Fourier(1024), fm.Fourier(1024))
        rec('kron_fourier_1024x1024_c64', timed(lambda: K.forward(x), k2, 3), cols, 16.0 * N)
        Hd = fm.Hadamard(20)
        xf = torch.view_as_real(x.t().contiguous()).reshape(cols * 2, N).t()             # (2^20, 2048) float32, column-major
        rec('hadamard_forward_o20_f32', timed(lambda: Hd.forward(xf), k2, 3), xf.shape[1], 8.0 * N)
        try:            # scatter -> FWHT(order 20) -> gather, three launches; 2 x 4 B x (2^20 - 1) per column
            Ll = fm.LFSRCirculant((1 << 20) | (1 << 3) | 1, 1)
            xl = xf[:N - 1, :1024].t().contiguous().t()
            rec('lfsr_circulant_forward_o20_f32', timed(lambda: Ll.forward(xl), k2, 3), xl.shape[1], 8.0 * (N - 1))
            del xl, Ll
        except Exception as e:                                     # a secondary row must not take the headline down
            extras['lfsr_circulant_forward_o20_f32'] = {'error': repr(e)}
        del xf
        Fb = fm.Fourier(1000003)
        xb = x[:1000003, :256].t().contiguous().t()
        rec('fourier_bluestein_1000003_c64', timed(lambda: Fb.forward(xb), 3, 2), 256, 16.0 * 1000003)
        del xb
        x16 = crandn(1 << 16, 64).to(torch.complex128)
        F16 = fm.Fourier(1 << 16)
        rec('fourier_forward_2^16_c128_64cols', timed(lambda: F16.forward(x16), 10, 3), 64, 32.0 * (1 << 16))

    cpu = None
    if rank == 0 and not args.no_cpu and not distributed:          # reported at N = 1 only
        cpu = cpu_baseline_single(cols=args.cpu_cols, reps=2)

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
            'e2e': e2e,
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_step': alg_bytes,
                         'note': ('operator level: algorithmic bytes of one step (2 x 8 B x N per column, SURVEY 8d) / CUDA-event '
                                  'time of the step on the caller stream; one step = %d launches (3 passes per slab of %d columns, '
                                  'slabs issued round-robin on internal streams so a per-kernel event time does not exist); '
                                  'kernels = each pass kernel\'s share of the step and traffic = DRAM bytes per step, both from '
                                  'the committed ncu launch list of this command (profiles/r1_traffic.json)')
                                 % (launches // steps, int(info.slab_cols)),
                         'dominant_kernel': dominant,
                         'launches_per_step': launches // steps,
                         'kernels': kernels},
            'cpu_baseline': cpu,
            'extras': extras,
        }
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cols', type=int, default=COLS)
    ap.add_argument('--e2e-cols', type=int, default=256, dest='e2e_cols')
    ap.add_argument('--cpu-cols', type=int, default=16, dest='cpu_cols')
    ap.add_argument('--quick', action='store_true', help='headline only (no extras)')
    ap.add_argument('--no-e2e', action='store_true', dest='no_e2e')
    ap.add_argument('--no-cpu', action='store_true', dest='no_cpu')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
