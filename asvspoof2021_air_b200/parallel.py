"""Data parallelism: one process per GPU (torchrun), NCCL over NVLink/NVSwitch for the single exchange
step of the path -- the sum all-reduce of the flat gradient buffer.

The reference is single-GPU (a commented-out nn.DataParallel at main_train.py:174); utterances are
independent through LFCC / crop-pad / conv stacks / pooling, BatchNorm statistics stay per replica (the
reference has no SyncBN, so a per-GPU batch of B reproduces its batch-B semantics on every replica) and
the loss is a per-replica mean.  The 1/world average is folded into the optimiser kernels (grad_scale).

`GradReducer` launches the all-reduce in buckets on a side stream as the backward pass finishes the
layers that own them (the flat buffer is in forward order, the backward pass walks it from the end), so
the exchange overlaps the remaining dgrad/wgrad work; the optimiser waits for the last bucket.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank)."""
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # the ring / NVLS kernels of the gradient exchange run BESIDE persistent 148-CTA conv grids: a handful of
            # CTAs moves 50 MB per step well inside the backward pass, more only evicts compute (override via the env)
            # (Trainer leaves exactly that many SMs out of its persistent grids during the backward pass)
            os.environ.setdefault("NCCL_MAX_CTAS", "2")
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local if world > 1 else torch.cuda.current_device())
    return rank, world, local


def collective_ctas():
    """CTAs (= SMs) the gradient exchange may occupy while compute kernels are running (the NCCL_MAX_CTAS cap)."""
    try:
        return max(0, int(os.environ.get("NCCL_MAX_CTAS", "0")))
    except ValueError:
        return 0


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n utterances for this rank (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bucket_bounds(n, bucket_elems, cuts=()):
    """Bucket boundaries of a flat buffer of n elements, walked from the END (backward order):
    [(lo, hi), ...] with hi descending.  `cuts`: offsets that must be bucket boundaries (the offsets at which the
    backward pass reports "everything above is final": a bucket straddling one would wait for the next report)."""
    out, hi = [], n
    cuts = sorted({int(c) for c in cuts if 0 < c < n}, reverse=True)
    while hi > 0:
        lo = max(0, hi - bucket_elems)
        for c in cuts:
            if lo < c < hi:
                lo = c
                break
        out.append((lo, hi))
        hi = lo
    return out


class GradReducer:
    """Bucketed sum all-reduce of a flat gradient buffer, overlapped with the backward pass."""

    def __init__(self, flat, n, group=None, bucket_elems=4 << 20, overlap=True, cuts=()):
        self.flat, self.n, self.group = flat, n, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = bucket_bounds(n, bucket_elems, cuts)
        self.overlap = overlap and flat.is_cuda
        self.stream = torch.cuda.Stream() if self.overlap else None
        self.next = 0
        self.extra = []
        self.trace = [] if os.environ.get("AIR_DDP_TRACE") == "1" else None

    def begin(self):
        self.next = 0

    def ready(self, offset, wait_events=()):
        """Every gradient at flat index >= offset is final once the current stream AND `wait_events` (e.g. the
        weight-gradient side stream of the engine) have reached this point: launch the buckets that lie above it.  The
        exchange stream waits for those events; the compute stream does not wait for anything."""
        if self.world == 1:
            return
        while self.next < len(self.buckets) and self.buckets[self.next][0] >= offset:
            self._launch(self.buckets[self.next], wait_events)
            self.next += 1

    def _launch(self, rng, wait_events=()):
        lo, hi = rng
        view = self.flat[lo:hi]
        if self.overlap:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                for e in wait_events:
                    self.stream.wait_event(e)
                dist.all_reduce(view, group=self.group)
        else:
            for e in wait_events:
                e.synchronize()
            dist.all_reduce(view, group=self.group)

    def finish(self, *extra):
        """Launch whatever is left (and the small extra tensors, e.g. the OC-Softmax centre gradient), then make
        the compute stream wait for the exchange.  Returns the factor the optimiser applies (1/world)."""
        if self.world == 1:
            return 1.0
        self.ready(0)
        for t in extra:
            if self.overlap:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(ev)
                    dist.all_reduce(t, group=self.group)
            else:
                dist.all_reduce(t, group=self.group)
        if self.overlap:
            if self.trace is not None:                       # AIR_DDP_TRACE=1: how long the compute stream waits for the exchange
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record(torch.cuda.current_stream())       # backward pass done on the compute stream
                e1.record(self.stream)                       # exchange done
                torch.cuda.current_stream().wait_stream(self.stream)
                e2.record(torch.cuda.current_stream())
                self.trace.append((e0, e2))
            else:
                torch.cuda.current_stream().wait_stream(self.stream)
        return 1.0 / self.world

    def exposed_ms(self):
        """Mean time per step the compute stream spent waiting for the gradient exchange (AIR_DDP_TRACE=1), else None."""
        if not self.trace:
            return None
        torch.cuda.synchronize()
        v = [a.elapsed_time(b) for a, b in self.trace[len(self.trace) // 2:]]
        return sum(v) / len(v)


def broadcast_state(tensors, src=0, group=None):
    """Make every replica start from rank `src`'s parameters / buffers."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)


def gather_ragged(columns, group=None):
    """All-gather per-rank 1-D tensors of different lengths (validation scores + labels of each rank's shard of the
    dev set): `columns` is a list of equally long 1-D tensors; returns the list of their concatenations over ranks,
    in rank order.  One size exchange + one padded all-gather; identity when not distributed."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return list(columns)
    world = dist.get_world_size(group)
    dev = columns[0].device
    n = torch.tensor([columns[0].numel()], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x) for x in sizes]
    pad = torch.zeros(len(columns), max(max(sizes), 1), device=dev, dtype=torch.float64)
    for i, c in enumerate(columns):
        pad[i, :c.numel()] = c.to(torch.float64)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [torch.cat([b[i, :k] for b, k in zip(bufs, sizes)]).to(c.dtype) for i, c in enumerate(columns)]
