"""Data parallelism over the batch: one process per GPU, torch.distributed (NCCL over NVLink) for the
plumbing.  The reference has no distributed code at all; what has to be exchanged for
an N-GPU run to equal the reference at the global batch size is exactly
  * BatchNorm batch statistics, forward  [sum x, sum x^2]    2C fp32 per layer per pass   (SyncBN)
  * BatchNorm backward reductions        [sum g, sum g*xhat] 2C fp32 per layer per pass
  * parameter gradients (sum over the global batch)           one flat bucket per network
  * logged scalars (means over the global batch).
Rank r owns rows [r*B/W, (r+1)*B/W) of every per-sample tensor; parameters, Adam state and BN buffers
are replicated and stay identical because every rank applies the same update to the same reduced
gradient."""
import os

import torch
import torch.distributed as dist


_PEER_MAX_N = 3072       # JCK_COMM_MAX_N


class LocalComm:
    """World of one: every collective is the identity."""
    world_size = 1
    rank = 0
    peer = None
    peer_aux = None
    transport = "local"

    def allreduce_sum_(self, t):
        return t

    def allreduce_mean_(self, t):
        return t

    def allreduce_mean_begin(self, t):
        return None

    def allreduce_mean_end(self, handle, t):
        return t

    def barrier(self):
        pass

    def check_health(self):
        pass


class TorchComm:
    """Collectives through an initialised torch.distributed process group (nccl on GPUs, gloo in the
    CPU tests)."""

    peer = None      # opaque libjck_b200 communicator (NVLink peer-memory mailboxes) once open_peer() succeeded
    peer_aux = None  # a second, independent one (own mailboxes and call counter) for exchanges issued from a second stream

    def __init__(self, group=None):
        assert dist.is_initialized(), "init the process group first (see init_from_env)"
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def open_peer(self):
        """Open the peer-memory communicators of csrc/comm.cu (the main one and the auxiliary one)."""
        if self.peer is not None or self.world_size == 1 or self.world_size > 8 or not torch.cuda.is_available():
            return self
        self.peer = self._open_one()
        if self.peer is not None:
            self.peer_aux = self._open_one()       # for exchanges issued from a second stream (AuxComm)
        return self

    @property
    def transport(self):
        """What carries the SyncBN statistics: 'peer' (NVLink peer-memory mailboxes, csrc/comm.cu) or 'nccl'."""
        return "peer" if self.peer is not None else "nccl"

    def _open_one(self):
        """One peer-memory communicator: every rank exports its mailbox with CUDA IPC, the
        handles travel through one torch.distributed all_gather, every rank maps its peers.  From then on the
        SyncBN exchanges (2C floats, 40+ per step) are single-CTA kernels storing straight into the peers' HBM
        instead of host-launched NCCL calls.  Large buffers (the gradient buckets) stay on NCCL.
        Returns the opaque handle, or None (with a warning on rank 0) when any rank could not map any peer."""
        from . import ops
        dev = torch.device("cuda", torch.cuda.current_device())
        comm, err = None, ""
        try:
            comm, handle = ops.comm_create(self.rank, self.world_size)
            mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
        except Exception as e:          # noqa: BLE001 -- e.g. CUDA IPC not permitted in this container
            err, mine = str(e), torch.zeros(ops.COMM_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world_size)]
        dist.all_gather(gathered, mine, group=self.group)
        if comm is not None:
            try:
                ops.comm_connect(comm, b"".join(bytes(g.cpu().tolist()) for g in gathered))
            except Exception as e:      # noqa: BLE001 -- no peer access between two of the GPUs
                err = str(e)
        # every rank must take the same transport: the mailboxes are used only if ALL ranks mapped ALL peers
        ok = torch.tensor([0 if err else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)   # also: every mailbox is zeroed and mapped from here on
        if int(ok) == 0:
            if self.rank == 0:
                import warnings
                warnings.warn("peer-memory communicator unavailable (%s); SyncBN statistics travel over NCCL" % (err or "a peer failed"))
            if comm is not None:
                ops.comm_destroy(comm)
            return None
        return comm

    def allreduce_sum_(self, t):
        if self.world_size > 1:
            if self.peer is not None and t.is_cuda and t.dtype == torch.float32 and t.numel() <= _PEER_MAX_N \
                    and t.is_contiguous():
                from . import ops
                ops.comm_allreduce_small(self.peer, t)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_mean_(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world_size)
        return t

    def allreduce_mean_begin(self, t):
        """Start averaging a gradient bucket; kernels launched before allreduce_mean_end() overlap the exchange
        (the collective runs on the process group's own stream, ordered after everything queued so far)."""
        if self.world_size == 1:
            return None
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def allreduce_mean_end(self, handle, t):
        if handle is not None:
            handle.wait()                       # stream-ordered: the current stream waits, the host does not
            t.div_(self.world_size)
        return t

    def barrier(self):
        if self.world_size > 1:
            dist.barrier(group=self.group)

    def check_health(self):
        """Raise if a peer-memory exchange ever timed out (csrc/comm.cu: the exchange then returned NaN and raised a sticky
        flag instead of trapping).  Synchronises with the device: call it where the host syncs anyway (the trainers do,
        every 100 steps).  JCK_SYNCBN=nccl selects the NCCL transport, whose watchdog handles stragglers itself."""
        from . import ops
        for tag, c in (("main", self.peer), ("auxiliary", self.peer_aux)):
            if c is not None and ops.comm_error(c):
                raise RuntimeError(f"rank {self.rank}: a SyncBN exchange on the {tag} peer-memory communicator timed out waiting "
                                   "for a peer (JCK_COMM_TIMEOUT_S, default 120 s); the step's statistics were poisoned with NaN")


class AuxComm:
    """The same communicator as seen from a SECOND stream: its small (SyncBN) exchanges go through the auxiliary
    peer communicator, whose mailboxes and device-side call counter are its own, so they cannot interleave with the main
    stream's exchanges (every rank issues each stream's exchanges in the same order, but the two streams' relative order
    on the device is free).  Large or NCCL-carried exchanges fall through to the process group, which serialises them in
    host issue order -- identical on every rank."""

    def __init__(self, comm):
        self._c = comm
        self.world_size, self.rank = comm.world_size, comm.rank
        self.peer = getattr(comm, "peer_aux", None)
        self.peer_aux = None
        self.transport = getattr(comm, "transport", "local")

    def allreduce_sum_(self, t):
        if self.world_size > 1:
            if self.peer is not None and t.is_cuda and t.dtype == torch.float32 and t.numel() <= _PEER_MAX_N \
                    and t.is_contiguous():
                from . import ops
                ops.comm_allreduce_small(self.peer, t)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._c.group)
        return t

    def barrier(self):
        self._c.barrier()


def init_from_env(backend=None):
    """Build the communicator torchrun's environment describes (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT); world of one when launched plainly."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return LocalComm()
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    comm = TorchComm()
    if dist.get_backend() == "nccl" and os.environ.get("JCK_SYNCBN", "p2p") == "p2p":
        comm.open_peer()
    return comm


def env_rank_world():
    """(rank, world size) torchrun's environment describes; (0, 1) when launched plainly"""
    return int(os.environ.get("RANK", "0")), max(1, int(os.environ.get("WORLD_SIZE", "1")))


def local_slice(n, rank, world):
    """Rows of an n-row GLOBAL batch that rank `rank` of `world` trains on: n // world consecutive rows (rank r
    owns rows [r B/W, (r+1) B/W)); when world does not divide n the last n % world rows are dropped on every rank so that all
    ranks keep the same local batch (SyncBN's sample count is static)."""
    per = n // world
    return slice(rank * per, (rank + 1) * per)


def shard_rows(t, comm):
    """Rows of a global per-sample tensor owned by this rank."""
    return t[local_slice(t.shape[0], comm.rank, comm.world_size)]


class FlatParams:
    """Re-home a module's parameters, gradients and Adam moments into flat fp32 buffers (views keep the
    nn.Parameter API intact): Adam becomes one launch per network and the gradient exchange one
    all-reduce per network."""

    ALIGN = 32       # floats

    def __init__(self, module):
        params = [p for p in module.parameters()]
        self.params = params
        # every parameter starts on a 128-byte boundary (the kernels read parameters and gradients with 128-bit
        # accesses); the padding elements stay zero in all four buffers
        n = sum(-(-p.numel() // self.ALIGN) * self.ALIGN for p in params)
        dev = params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        # The flat buffers are laid out in REVERSE parameter order: backward produces the gradients of the last layer
        # first, so the gradients that become final together are contiguous and a bucket of them (GradBuckets) is one
        # slice.  `offsets` stays indexed by parameter order (optimizer state / checkpoints are unaffected).
        self.offsets = [None] * len(params)
        self.layout = list(range(len(params)))[::-1]
        off = 0
        for idx in self.layout:
            p = params[idx]
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            self.offsets[idx] = (off, k)
            off += -(-k // self.ALIGN) * self.ALIGN
        self.numel = n

    def views(self, flat):
        return [flat[o:o + k].view(p.shape) for (o, k), p in zip(self.offsets, self.params)]

    def rebind(self):
        """Re-attach .grad views (model.zero_grad() sets them to None by default)."""
        for (o, k), p in zip(self.offsets, self.params):
            if p.grad is None or p.grad.data_ptr() != self.grad[o:o + k].data_ptr():
                p.grad = self.grad[o:o + k].view(p.shape)


class GradBuckets:
    """Gradient exchange of one network, bucketed and overlapped with the backward sweep that produces it
    (SURVEY.md 8e): the flat gradient buffer (reverse parameter order = the order backward finishes them) is cut into
    contiguous buckets of at least `min_elems` floats; the sweep calls ready(p) as each parameter's gradient becomes
    final, and the moment a bucket is complete its all-reduce (SUM -- the loss head already divides by the GLOBAL batch,
    jck_head_bwd mean_count) starts on the process group's stream while the sweep goes on.  finish() makes the
    current stream wait for all of them.  ready() may be called from any stream (the weight-gradient kernels run on a
    side stream): the bucket's collective is ordered after an event recorded at each call.
    Each collective is issued from a dedicated ISSUE stream that waits for exactly the events of its bucket and nothing
    else.  (Round-2 timeline at two GPUs, profiles/r02_scale_timeline.md: issued from whatever stream was current, the
    collective was ordered behind everything already queued there -- a bucket completed by conv3's weight gradient waited
    for conv2's, queued after it on the same side stream, so all three generator buckets ran back to back AFTER the last
    weight gradient, 100 us in front of Adam; issued from the home stream it made the sweep itself wait for the lagging
    weight-gradient stream.)  Issue order = completion order in host program order, identical on every rank."""

    def __init__(self, flat, comm, min_elems=1 << 19, merge_tail=None):
        self.flat, self.comm = flat, comm
        if merge_tail is None:
            merge_tail = os.environ.get("JCK_MERGE_TAIL", "1") != "0"
        self.buckets = []                # [lo, hi, {param ids}]
        lo, ids = 0, set()
        for idx in flat.layout:
            o, k = flat.offsets[idx]
            end = o + -(-k // flat.ALIGN) * flat.ALIGN          # padded end = the next parameter's offset
            ids.add(id(flat.params[idx]))
            if end - lo >= min_elems:
                self.buckets.append([lo, end, ids])
                lo, ids = end, set()
        if ids:
            self.buckets.append([lo, flat.numel, ids])
        # The last buckets complete when the sweep is over: nothing is left to hide them behind, they run back to back in
        # front of Adam and each pays the collective's latency (8-GPU timeline, profiles/r02_scale_timeline.md: the
        # generator's conv2.w and conv1.w buckets 82 + 51 us).  One collective for the tail instead of two.
        if merge_tail and len(self.buckets) >= 3:
            lo2, _, ids2 = self.buckets[-2]
            _, hi3, ids3 = self.buckets[-1]
            self.buckets[-2:] = [[lo2, hi3, ids2 | ids3]]
        self.bucket_of = {pid: b for b, (_, _, ids) in enumerate(self.buckets) for pid in ids}
        self.issue_stream = torch.cuda.Stream(device=flat.grad.device) if (flat.grad.is_cuda and comm.world_size > 1) else None
        self.begin()

    def begin(self):
        self.missing = [set(ids) for _, _, ids in self.buckets]
        self.events = [[] for _ in self.buckets]
        self.issued = [False] * len(self.buckets)
        self.handles = []

    def ready(self, *params):
        if self.comm.world_size == 1:
            return
        for p in params:
            b = self.bucket_of[id(p)]
            if self.issued[b]:
                continue
            ev = torch.cuda.Event() if p.is_cuda else None
            if ev is not None:
                ev.record()
                self.events[b].append(ev)
            self.missing[b].discard(id(p))
            if not self.missing[b]:
                self._issue(b)

    def _issue(self, b):
        if self.issued[b]:
            return
        lo, hi, _ = self.buckets[b]
        grp = getattr(self.comm, "group", None)
        if self.issue_stream is None:                      # CPU tensors (gloo tests)
            self.handles.append(dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM, group=grp, async_op=True))
        else:
            s = self.issue_stream
            if self.events[b]:
                for ev in self.events[b]:
                    s.wait_event(ev)
            else:                                          # never marked ready: ordered after the current stream
                s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):                     # the process group's stream orders itself after `s` only
                self.handles.append(dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.SUM, group=grp, async_op=True))
        self.issued[b] = True

    def finish(self):
        """Start whatever was never marked ready (ordered after the current stream), then wait for everything."""
        if self.comm.world_size == 1:
            return
        for b in range(len(self.buckets)):
            if not self.issued[b]:
                self._issue(b)
        for h in self.handles:
            h.wait()                      # stream-ordered: the current stream waits, the host does not
        if self.issue_stream is not None and self.handles:
            torch.cuda.current_stream().wait_stream(self.issue_stream)      # rejoin (CUDA-graph capture needs every fork joined)
        self.handles = []
