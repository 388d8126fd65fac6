"""A/B sweeps with the stand-alone C harness (build/cbench): one short process per (library, environment, operator).

    python tools/sweep_cbench.py <plan> [cols]

Plans are lists of (label, lib, env, op) defined below; every line of output is `label | <cbench line>`.  No torch import:
a configuration costs about a second.  Experiments only - the defaults of the library are the measured best."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastmat_b200", "lib", "libfastmat_b200.so")
R1 = os.path.join(ROOT, "build", "alt", "lib_r1.so")
CB = os.path.join(ROOT, "build", "cbench")


def run(label, lib, env, op, cols, reps=5, seconds=0):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    try:
        out = subprocess.run([CB, lib, op, str(cols), str(reps), str(seconds)], env=e, capture_output=True, text=True, timeout=300)
        line = (out.stdout.strip() or out.stderr.strip()).splitlines()[-1] if (out.stdout.strip() or out.stderr.strip()) else "no output rc=%d" % out.returncode
    except subprocess.TimeoutExpired:
        line = "TIMEOUT"
    print("%-44s | %s" % (label, line), flush=True)
    try:
        return float(line.split("mean")[1].split("ms")[0])
    except Exception:
        return float("inf")


def envs(d):
    return " ".join("%s=%s" % (k.replace("FMB_", ""), v) for k, v in d.items()) or "default"


def plan_kernels(cols):
    if os.path.exists(R1):
        for op in ("circ", "fourier", "toep"):
            run("r1 " + op, R1, {}, op, cols)
        run("r1 fourier V32P=0", R1, {"FMB_V32P": 0}, "fourier", cols)
    best = (float("inf"), None)
    for twm in (1, 0):
        for occ in (0, 1):
            for msh in (0, 1, 2, 3, 4):
                env = {"FMB_V32_TWM": twm, "FMB_V32_OCC": occ, "FMB_V32_MSHAPE": msh}
                t = run("circ " + envs(env), LIB, env, "circ", cols)
                if t < best[0]:
                    best = (t, env)
    print("best circ:", best, flush=True)
    benv = best[1]
    for mb in (8, 16, 24, 32, 48):
        for ns in (2, 3, 4, 6):
            env = dict(benv, FMB_PIPE_MB=mb, FMB_PIPE_STREAMS=ns)
            run("circ " + envs(env), LIB, env, "circ", cols)
    for occ in (0, 1):
        for v32p in (0, 1):
            env = {"FMB_V32_OCC": occ, "FMB_V32P": v32p}
            run("fourier " + envs(env), LIB, env, "fourier", cols)
            run("kron " + envs(env), LIB, env, "kron", cols)
    for occ in (0, 1):
        for msh in (0, 3, 4):
            env = {"FMB_V32_OCC": occ, "FMB_V32_MSHAPE": msh}
            run("toep " + envs(env), LIB, env, "toep", cols)
            run("toepb " + envs(env), LIB, env, "toepb", cols)


def plan_final(cols):
    for op in ("circ", "circb", "fourier", "toep", "toepb", "kron"):
        run("final " + op, LIB, {}, op, cols, reps=10, seconds=2)


if __name__ == "__main__":
    plan = sys.argv[1] if len(sys.argv) > 1 else "kernels"
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    {"kernels": plan_kernels, "final": plan_final}[plan](cols)
