"""V32P (persistent, TMA-fed) against the per-pass V32 kernels: bit-identical outputs + timings.

    python tools/check_v32p.py            # parent: runs itself once per configuration, compares the saved outputs
    python tools/check_v32p.py child OUT  # child: computes, saves the outputs to OUT, prints timings

The two paths execute the same butterflies in the same order per column, so their results must be bit-identical.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(out, cols):
    import numpy as np
    import torch
    import fastmat_b200 as fm
    N = 1 << 20
    rng = np.random.default_rng(0)
    g = torch.Generator(device='cuda').manual_seed(11)
    c = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)

    def crandn(m, n):
        return torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()

    res = {}
    xs = crandn(13, N)                                   # ragged for 2-column slabs
    C, F = fm.Circulant(c), fm.Fourier(N)
    nt = N // 2
    T = fm.Toeplitz(c[:nt].copy(), c[nt:2 * nt - 1].copy())
    xt = crandn(7, nt)
    for name, fn, x in (('circ_f', C.forward, xs), ('circ_b', C.backward, xs), ('four_f', F.forward, xs), ('four_b', F.backward, xs),
                        ('toep_f', T.forward, xt), ('toep_b', T.backward, xt), ('circ_1', C.forward, xs[:, 3].contiguous())):
        t0 = time.time()
        y = fn(x)
        torch.cuda.synchronize()
        res[name] = y.cpu()
        print('  %s done in %.2f s' % (name, time.time() - t0), flush=True)
    torch.save(res, out)
    del res

    def timed(f, k=5):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    x = crandn(cols, N)
    xt = crandn(cols, nt)
    gb = 16.0 * N * cols / 1e9
    peak = 6550.1
    tc, tb, tf, tt = timed(lambda: C.forward(x)), timed(lambda: C.backward(x)), timed(lambda: F.forward(x)), timed(lambda: T.forward(xt))
    print('  TIMING %d cols | circ fwd %.3f ms (%.1f%%) bwd %.3f ms | fourier %.3f ms (%.1f%%) | toeplitz %.3f ms (%.1f%% of s(m+n))' % (
        cols, tc, gb / tc * 1e3 / peak * 100, tb, tf, gb / tf * 1e3 / peak * 100, tt, gb / 2 / tt * 1e3 / peak * 100), flush=True)


def main():
    import torch
    cols = int(os.environ.get('CHECK_COLS', '512'))
    on = {'FMB_V32P': '2'}                 # persistent path for convolutions too (default: plain transforms only)
    configs = [('v32 per-pass (reference)', {'FMB_V32P': '0'}),
               ('v32p defaults (conv: slab 2; plain: slab 1; blocked order)', dict(on)),
               ('v32p slab=1', dict(on, FMB_V32P_SLAB='1')),
               ('v32p slab=2', dict(on, FMB_V32P_SLAB='2')),
               ('v32p slab=3', dict(on, FMB_V32P_SLAB='3')),
               ('v32p ahead=3', dict(on, FMB_V32P_AHEAD='3')),
               ('v32p ahead=8', dict(on, FMB_V32P_AHEAD='8')),
               ('v32p interleaved order', dict(on, FMB_V32P_MIX='1')),
               ('v32p no L2 hints', dict(on, FMB_V32P_HINTS='0')),
               ('v32p L2 promotion 128 B', dict(on, FMB_V32P_PROMO='2'))]
    extra = os.environ.get('CHECK_ONLY')
    if extra:
        configs = [configs[0]] + [c for c in configs[1:] if extra in c[0]]
    ref = None
    bad = 0
    hangs = 0
    for i, (name, env) in enumerate(configs):
        out = '/tmp/v32p_check_%d.pt' % i
        e = dict(os.environ)
        e.update(env)
        print('== %s' % name, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), 'child', out, str(cols)], env=e, cwd=ROOT,
                               capture_output=True, text=True, timeout=int(os.environ.get('CHECK_TIMEOUT', '120')))
        except subprocess.TimeoutExpired as ex:
            print('  TIMEOUT (hang?)', (ex.stdout or b'')[-2000:], flush=True)
            bad += 1
            hangs += 1
            if hangs >= 2:
                print('  two hangs: giving up', flush=True)
                break
            continue
        print(r.stdout.rstrip()[-3000:], flush=True)
        if r.returncode != 0:
            print('  FAILED rc=%d\n%s' % (r.returncode, r.stderr[-3000:]), flush=True)
            bad += 1
            continue
        res = torch.load(out)
        if ref is None:
            ref = res
            continue
        for k in ref:
            same = torch.equal(ref[k], res[k])
            if not same:
                d = (ref[k] - res[k]).abs()
                print('  MISMATCH %s: max abs diff %.3e, %d of %d elements differ, nan=%d' % (
                    k, float(d.max()), int((d > 0).sum()), d.numel(), int(torch.isnan(res[k].real).sum())), flush=True)
                if d.dim() == 2:
                    print('    per column:', [int(v) for v in (d > 0).sum(0)], flush=True)
                bad += 1
        print('  outputs bit-identical to the per-pass path' if not bad else '  (mismatches so far: %d)' % bad, flush=True)
    print('CHECK %s' % ('OK' if not bad else 'FAILED (%d)' % bad))
    return 1 if bad else 0


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child(sys.argv[2], int(sys.argv[3]))
    else:
        sys.exit(main())
