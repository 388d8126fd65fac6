"""Column-batch data parallelism for the apply path.

Every operator acts column-wise -- ``forward(x)[:, m]`` depends only on ``x[:, m]`` (e.g. fastmat/Hadamard.pyx:180,
fastmat/core/cmath.pyx:992-997, ``np.fft.fft(axis=0)``) -- so the M columns of a 2-D operand are split into contiguous
blocks, one per rank (one process per GPU, ``torch.distributed``), and each rank applies its own plan to its own block.
There is NO collective on the apply path.  Only when the caller wants the assembled result does ``gather_columns`` issue
one all_gather / gather (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).
"""
import torch
import torch.distributed as dist


def shard_bounds(num_cols, rank, world_size):
    """Contiguous, balanced column block of ``rank``: the first ``num_cols % world_size`` ranks get one extra column."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(num_cols), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_columns(x, rank=None, world_size=None):
    """View of this rank's column block of a full (n, M) operand (no copy)."""
    if rank is None:
        rank = dist.get_rank()
    if world_size is None:
        world_size = dist.get_world_size()
    if x.ndim != 2:
        raise ValueError("column sharding needs a 2D operand")
    a, b = shard_bounds(x.shape[1], rank, world_size)
    return x[:, a:b]


def apply_sharded(matrix, x_local, backward=False):
    """Apply ``matrix`` to this rank's column block.  No communication."""
    return matrix.backward(x_local) if backward else matrix.forward(x_local)


def share_largest_singular_value(matrix, src=0, group=None):
    """ISTA / FISTA step size: the power iteration runs on rank ``src`` only and the scalar is broadcast (8 bytes), so
    that every rank iterates with the identical step size; cached on the matrix like the property itself."""
    dev = matrix._default_device() if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    t = torch.zeros(1, dtype=torch.float64, device=dev)
    if dist.get_rank(group) == src:
        t[0] = float(matrix.largestSingularValue)
    dist.broadcast(t, src=src, group=group)
    matrix._cache['lsv'] = float(t.item())
    return matrix._cache['lsv']


def solve_sharded(algorithm, b_local, share_step_size=True):
    """Run a fastmat_b200.algorithms solver on this rank's block of right-hand sides (columns are independent problems:
    no collective inside the iteration).  With ``share_step_size`` the largest singular value is computed once and
    broadcast first (only ISTA / FISTA need it)."""
    if share_step_size and dist.is_initialized() and hasattr(algorithm, 'numLambda'):
        share_largest_singular_value(algorithm.fmatA)
    return algorithm.process(b_local)


def gather_columns(y_local, num_cols, dst=None, group=None):
    """Assemble the (n, num_cols) result from the per-rank column blocks.

    ``dst=None``: every rank gets the full array (all_gather); otherwise only rank ``dst`` does (others return None).
    Blocks may differ by one column (``shard_bounds``); they travel padded to the widest block.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    widths = [shard_bounds(num_cols, r, world)[1] - shard_bounds(num_cols, r, world)[0] for r in range(world)]
    if y_local.ndim != 2 or y_local.shape[1] != widths[rank]:
        raise ValueError("local block has %s columns, expected %d" % (tuple(y_local.shape), widths[rank]))
    wmax = max(widths)
    n = y_local.shape[0]
    # communicate column-major blocks: (wmax, n) contiguous buffers
    send = torch.zeros((wmax, n), dtype=y_local.dtype, device=y_local.device)
    send[:widths[rank]] = y_local.t()
    is_cplx = send.is_complex()
    if is_cplx:
        send = torch.view_as_real(send)         # the backends move real buffers (gloo has no complex types)
    if dst is None:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send, group=group)
    else:
        recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send, recv, dst=dst, group=group)
        if rank != dst:
            return None
    out = torch.empty((num_cols, n), dtype=y_local.dtype, device=y_local.device)
    c0 = 0
    for r in range(world):
        blk = torch.view_as_complex(recv[r]) if is_cplx else recv[r]
        out[c0:c0 + widths[r]] = blk[:widths[r]]
        c0 += widths[r]
    return out.t()                      # (n, num_cols), column-major like every result of the apply path


def bind_to_gpu_numa(device_index):
    """Restrict this process to the host cores NVML reports as local to GPU ``device_index`` (sched_setaffinity), so that
    pinned staging buffers allocated afterwards are first-touched on that GPU's NUMA node and the copy threads run next to
    its PCIe root.  Host-path (apply_host) scaling over the GPUs of one box depends on it when the box has several NUMA
    nodes; a no-op (returns None) when NVML or the affinity call is unavailable.  Returns the sorted core list."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and w * 64 + b < ncpu]
        if not cores:
            return None
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0))) or sorted(cores)
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def host_copy_ceiling(x_host, y_host, device, chunk_bytes=64 << 20, repeats=1):
    """Copy-only ceiling of the host path on this rank: the input batch host->device and an equally chunked output batch
    device->host, on two streams at once, nothing computed.  Returns (seconds, h2d_bytes, d2h_bytes) - what apply_host
    could reach if the transform were free."""
    import time
    rows, M = x_host.shape
    step = max(1, min(M, chunk_bytes // max(1, rows * x_host.element_size())))
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    col_major = not (x_host.shape[1] > 1 and x_host.stride(1) == 1 and x_host.stride(0) != 1)
    def dev_buf(r, c, dt):
        return torch.empty((c, r), dtype=dt, device=device).t() if col_major else torch.empty((r, c), dtype=dt, device=device)
    din = [dev_buf(rows, step, x_host.dtype) for _ in range(2)]
    dout = [dev_buf(y_host.shape[0], step, y_host.dtype) for _ in range(2)]
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(repeats):
        for k, c0 in enumerate(range(0, M, step)):
            c1 = min(M, c0 + step)
            with torch.cuda.stream(s_in):
                din[k % 2][:, :c1 - c0].copy_(x_host[:, c0:c1], non_blocking=True)
            with torch.cuda.stream(s_out):
                y_host[:, c0:c1].copy_(dout[k % 2][:, :c1 - c0], non_blocking=True)
    s_in.synchronize()
    s_out.synchronize()
    dt = (time.perf_counter() - t0) / repeats
    return dt, x_host.numel() * x_host.element_size(), y_host.numel() * y_host.element_size()
