"""Multi-GPU decomposition of the flame-front path (one process per GPU, torch.distributed).

Frames shard naturally:

* within a clip: contiguous frame ranges, one per rank (the contiguous variant of the
  reference's ``distribute_indices``, src/photron/parallel.py:101-113), each prefixed by a
  one-frame halo read from host memory - the only cross-frame dependency is the one-frame
  look-back of the frame difference (scripts/process_videos.py:397-399,469);
* across clips: whole videos round-robin over ranks (src/photron/parallel.py:173-208).

The single real exchange is tiny - the first exit frame (one int32, min over ranks) and the
per-rank position / count arrays - and it is ONE step: every rank's range kernel writes into
a *range block* ``{first_exit,0,0,0 | pos[cap] | counts[cap]}`` and one kernel per rank turns
the blocks of all ranks into the truncated whole-clip arrays (csrc/ff_exchange.cu).  It
replaces the reference's pickled ``comm.gather`` (scripts/process_videos.py:1533-1541) and
turns the per-rank ``break`` (:1494) into a global truncation (README.md:145-149).

Transports for the blocks:

``peer``      (GPUs of one box) blocks stay in place; peers map every rank's exchange buffer over
              NVLink through CUDA IPC.  The range kernel's last CTA publishes the block (epoch flag
              in peer memory), the merge kernel waits for the flags, pulls the blocks and acknowledges
              - it runs on a SIDE stream, off the critical path of the next clip - and every exit
              frame is min-reduced into an exit word on every rank, which lets the host-streamed path
              stop uploading behind an exit another rank found.  No collective on the data path.
``gathered``  one ``all_gather_into_tensor`` of the blocks (NCCL; gloo in the CPU tests), then
              ``ff_merge_ranges``.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ._cabi import FF_NO_EXIT, RangeHooks

BLOCK_HEADER = 4        # int32 words in front of a range block's arrays (csrc/ff_exchange.cu: kHdr)


def contiguous_range(total: int, rank: int, size: int) -> Tuple[int, int]:
    """[start, stop) of rank's block; the first ``total % size`` ranks get one extra item."""
    if size <= 0 or not 0 <= rank < size:
        raise ValueError(f"bad rank/size {rank}/{size}")
    base, extra = divmod(max(0, total), size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def assign_videos(n_videos: int, rank: int, size: int, weights: Optional[Sequence[int]] = None) -> List[int]:
    """Videos handled by ``rank``.  Default: round-robin like ``process_videos``
    (src/photron/parallel.py:192).  With ``weights`` (e.g. bytes per video): greedy
    longest-first balancing, deterministic on every rank."""
    if weights is None:
        return list(range(rank, n_videos, size))
    if len(weights) != n_videos:
        raise ValueError("weights must have one entry per video")
    loads = [0] * size
    owner = [0] * n_videos
    for v in sorted(range(n_videos), key=lambda i: (-weights[i], i)):
        r = min(range(size), key=lambda k: (loads[k], k))
        owner[v] = r
        loads[r] += weights[v]
    return [v for v in range(n_videos) if owner[v] == rank]


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pin this process to the CPUs local to its GPU (``/sys/bus/pci/devices/<bdf>/local_cpulist``)
    so that the pinned staging memory it allocates afterwards is first-touched on the GPU's own
    NUMA node: with 8 ranks streaming ~55 GB/s each, host memory placed on the wrong socket
    becomes the bottleneck.  Returns the CPU list, or None when the topology is not exposed."""
    bdf = None
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            raw = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            raw = (raw.decode() if isinstance(raw, bytes) else str(raw)).lower()
            bdf = raw[4:] if len(raw.split(":")[0]) == 8 else raw      # 00000000:1b:00.0 -> 0000:1b:00.0
        except Exception:
            return None
    path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
    try:
        text = open(path).read().strip()
        cpus: List[int] = []
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.extend(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


@dataclass
class RangeBlock:
    """This rank's outputs for one step, laid out for the exchange."""
    total: int                      # frames in the whole clip
    cap: int                        # capacity of pos / counts (ceil(total / world) or larger)
    pos: torch.Tensor               # int32[cap] view  - ff_detect's pos_out
    counts: torch.Tensor            # int32[cap] view  - ff_detect's count_out
    first_exit: torch.Tensor        # int32[1] view    - ff_detect's first_exit
    send: Optional[torch.Tensor] = None     # gathered transport: the whole block
    hooks: object = None            # peer transport: _cabi.RangeHooks for ff_process_range / ff_process_host_range
    init_first_exit: bool = False   # peer transport: the prep kernel resets first_exit (after waiting for the peers)


class GatheredRange:
    """The whole clip's results on this rank.  With the peer transport they are produced by the merge
    kernel on a side stream: reading an attribute makes the current stream wait for it."""

    def __init__(self, first_exit_t: torch.Tensor, pos: torch.Tensor, counts: Optional[torch.Tensor],
                 ready: Optional["torch.cuda.Event"] = None, check=None):
        self._first_exit_t, self._pos, self._counts, self._ready, self._check = first_exit_t, pos, counts, ready, check

    def wait(self) -> "GatheredRange":
        if self._ready is not None:
            torch.cuda.current_stream(self._pos.device).wait_event(self._ready)
            self._ready = None
        return self

    @property
    def first_exit_t(self) -> torch.Tensor:      # int32[1] global index of the first exit frame, or FF_NO_EXIT
        return self.wait()._first_exit_t

    @property
    def pos(self) -> torch.Tensor:               # int32[total] positions of the whole clip, truncated
        return self.wait()._pos

    @property
    def counts(self) -> Optional[torch.Tensor]:  # int32[total] above-noise counts
        return self.wait()._counts

    @property
    def first_exit(self) -> int:                 # synchronises on the device scalar (and on the exchange's status)
        value = int(self.first_exit_t.item())
        if self._check is not None:
            self._check()
        return value


class _DevicePointerView:
    """Expose raw device memory (owned by libflamefront) to torch through the CUDA array interface."""

    def __init__(self, ptr: int, n_int32: int):
        self.__cuda_array_interface__ = {"shape": (n_int32,), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


MergeFn = Callable[[torch.Tensor, int, int, int, torch.Tensor, Optional[torch.Tensor], torch.Tensor], None]


class RangeExchange:
    """The exit-min + truncation + result-gather step over a process group.

    ``engine``: the rank's ``FlameFrontEngine`` (supplies ``ff_merge_ranges`` and, for the peer
    transport, the C-ABI handle).  ``transport``: "auto" (peer memory when every rank is a GPU of
    this host and the IPC mapping succeeds, else gathered), "peer" or "gathered".  Without an
    engine (CPU tests of the host logic) the transport is "gathered" and ``finish`` needs a
    ``merge`` callable."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, engine=None, transport: str = "auto"):
        if transport not in ("auto", "peer", "gathered"):
            raise ValueError("transport must be 'auto', 'peer' or 'gathered'")
        self.group = group
        self.engine = engine
        self.active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.active else 0
        self.size = dist.get_world_size(group) if self.active else 1
        self._want = transport
        self.transport = "gathered"
        self._xchg = None               # ff_exchange* (peer transport)
        self._xchg_cap = 0
        self._xchg_buf = None           # torch view of the exchange's local allocation
        self._bufs = {}                 # gathered transport: (cap, device) -> (send, recv, init)
        self._peer_failed = False
        self._views = {}                # peer transport: torch views of the two block slots
        self._merge_stream = None       # peer transport: the merge kernels run here
        self._last_merge = None         # event behind the latest merge

    # ------------------------------------------------------------------ partition
    def my_range(self, total: int) -> Tuple[int, int]:
        return contiguous_range(total, self.rank, self.size)

    def block_cap(self, total: int) -> int:
        return max(1, -(-total // self.size))

    # ------------------------------------------------------------------ peer transport set-up
    def _peer_possible(self) -> bool:
        if self._want == "gathered" or self._peer_failed or self.engine is None:
            return False
        if not (self.active and self.size > 1):
            return False
        if dist.get_backend(self.group) != "nccl":
            return False
        return True

    def _setup_peer(self, cap: int) -> bool:
        """Create the exchange context and map the peers' blocks.  Collective: every rank calls
        it with the same ``cap``; the outcome is agreed on by all ranks."""
        eng = self.engine
        lib = eng._lib
        dev = eng.device
        ok = 1
        ctx = C.c_void_p()
        hb = lib.ff_exchange_handle_bytes()
        handle = (C.c_ubyte * hb)()
        try:
            # all ranks must be GPUs of one host
            names = [None] * self.size
            dist.all_gather_object(names, socket.gethostname(), group=self.group)
            if len(set(names)) != 1:
                ok = 0
            if ok:
                with torch.cuda.device(dev):
                    st = lib.ff_exchange_create(dev.index, self.rank, self.size, cap, C.byref(ctx))
                    if st == 0:
                        st = lib.ff_exchange_get_handle(ctx, handle)
                    if st != 0:
                        ok = 0
        except Exception:
            ok = 0
        mine = torch.tensor(list(bytes(handle)) + [ok], dtype=torch.uint8, device=dev)
        allh = torch.empty(self.size * (hb + 1), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        allh_np = allh.cpu().numpy().reshape(self.size, hb + 1)
        if not bool(allh_np[:, hb].all()):
            ok = 0
        if ok:
            packed = allh_np[:, :hb].copy().tobytes()
            with torch.cuda.device(dev):
                if lib.ff_exchange_open_peers(ctx, packed) != 0:
                    ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            if ctx.value:
                lib.ff_exchange_destroy(ctx)
            if self._want == "peer":
                raise RuntimeError("peer-memory exchange requested but CUDA IPC mapping of the peers failed")
            self._peer_failed = True
            return False
        self._xchg, self._xchg_cap = ctx, cap
        self.transport = "peer"
        return True

    def close(self) -> None:
        if self._xchg is not None:
            if self.active:
                torch.cuda.synchronize(self.engine.device)
                dist.barrier(group=self.group)       # no peer may still be reading my blocks
            self.engine._lib.ff_exchange_destroy(self._xchg)
            self._xchg = None
            self._xchg_buf = None
            self._views = {}
            self._last_merge = None

    # ------------------------------------------------------------------ one step
    def begin(self, total: int, device: Optional[torch.device] = None) -> RangeBlock:
        """Start a step: returns the block the range kernel fills
        (``engine.process_range(..., **exchange.range_kwargs(blk))``)."""
        cap = self.block_cap(total)
        if device is None:
            device = self.engine.device if self.engine is not None else torch.device("cpu")
        if self._peer_possible() and (self._xchg is None or cap > self._xchg_cap):
            if self._xchg is not None:
                self.close()
            self._setup_peer(cap)
        if self._xchg is not None:
            eng = self.engine
            pos_p, cnt_p, fe_p = C.c_void_p(), C.c_void_p(), C.c_void_p()
            hooks = RangeHooks()
            st = eng._lib.ff_exchange_begin(self._xchg, C.byref(pos_p), C.byref(cnt_p), C.byref(fe_p), C.byref(hooks))
            if st != 0:
                raise RuntimeError(f"ff_exchange_begin failed ({st})")
            xc = self._xchg_cap
            views = self._views.get((pos_p.value, xc))
            if views is None:

                def view(ptr, n):
                    return torch.as_tensor(_DevicePointerView(ptr.value, n), device=eng.device)
                views = self._views[(pos_p.value, xc)] = (view(pos_p, xc), view(cnt_p, xc), view(fe_p, 1))
            return RangeBlock(total, xc, views[0], views[1], views[2], None, hooks, True)
        key = (cap, str(device))
        if key not in self._bufs:
            n = BLOCK_HEADER + 2 * cap
            send = torch.zeros(n, dtype=torch.int32, device=device)
            recv = torch.zeros(self.size * n, dtype=torch.int32, device=device)
            init = torch.tensor([FF_NO_EXIT, 0, 0, 0], dtype=torch.int32, device=device)
            self._bufs[key] = (send, recv, init)
        send, _, init = self._bufs[key]
        send[:BLOCK_HEADER].copy_(init)
        return RangeBlock(total, cap, send[BLOCK_HEADER:BLOCK_HEADER + cap],
                          send[BLOCK_HEADER + cap:BLOCK_HEADER + 2 * cap], send[0:1], send)

    def finish(self, blk: RangeBlock, merge: Optional[MergeFn] = None, want_counts: bool = True) -> GatheredRange:
        """Exchange + merge: every rank returns the whole clip's truncated positions, counts and
        the global first exit frame.  ``merge`` defaults to the engine's ``ff_merge_ranges``.

        Peer transport: the block was published by the kernel that filled it (``hooks``); the merge kernel
        is launched on a side stream behind everything the current stream holds so far, and the returned
        ``GatheredRange`` makes its reader wait for it - the caller can go on with the next clip."""
        device = blk.pos.device
        if blk.send is None:            # peer transport
            eng = self.engine
            if self._merge_stream is None:
                self._merge_stream = torch.cuda.Stream(eng.device)
            side = self._merge_stream
            filled = torch.cuda.Event()
            filled.record(torch.cuda.current_stream(eng.device))
            side.wait_event(filled)
            with torch.cuda.stream(side):
                pos = torch.empty(blk.total, dtype=torch.int32, device=device)
                counts = torch.empty(blk.total, dtype=torch.int32, device=device) if want_counts else None
                fe = torch.empty(1, dtype=torch.int32, device=device)
                st = eng._lib.ff_exchange_finish(self._xchg, blk.total, pos.data_ptr(),
                                                 None if counts is None else counts.data_ptr(), fe.data_ptr(),
                                                 side.cuda_stream)
                if st != 0:
                    raise RuntimeError(f"ff_exchange_finish failed ({st})")
                ready = torch.cuda.Event()
                ready.record(side)
            main = torch.cuda.current_stream(eng.device)
            for t in (pos, counts, fe):
                if t is not None:
                    t.record_stream(main)
            self._last_merge = ready
            eng.launches += 1
            return GatheredRange(fe, pos, counts, ready, self.check)
        pos = torch.empty(blk.total, dtype=torch.int32, device=device)
        counts = torch.empty(blk.total, dtype=torch.int32, device=device) if want_counts else None
        fe = torch.empty(1, dtype=torch.int32, device=device)
        key = (blk.cap, str(device))
        send, recv, _ = self._bufs[key]
        if self.active and self.size > 1:
            dist.all_gather_into_tensor(recv, send, group=self.group)
        else:
            recv = send
        if merge is None:
            if self.engine is None:
                raise ValueError("finish needs an engine or a merge callable")
            merge = self.engine.merge_ranges
        merge(recv, self.size, blk.cap, blk.total, pos, counts, fe)
        return GatheredRange(fe, pos, counts)

    def join(self) -> None:
        """Make the current stream wait for the latest merge (peer transport; e.g. before a timing event)."""
        if self._last_merge is not None:
            torch.cuda.current_stream(self.engine.device).wait_event(self._last_merge)

    @staticmethod
    def range_kwargs(blk: RangeBlock) -> dict:
        """Keyword arguments that make ``engine.process_range`` fill ``blk`` in place (and, with the peer
        transport, wait for the peers first and publish the block from its last CTA)."""
        return dict(pos_out=blk.pos, counts_out=blk.counts, first_exit=blk.first_exit, truncate=False,
                    hooks=blk.hooks, init_first_exit=blk.init_first_exit)

    def acquire(self, blk: RangeBlock) -> None:
        """Before a block is filled by hand (peer transport): wait until no peer reads its previous use and
        reset its first-exit word - what the prep kernel does for ``engine.process_range``."""
        if blk.send is None:
            eng = self.engine
            st = eng._lib.ff_exchange_acquire(self._xchg, torch.cuda.current_stream(eng.device).cuda_stream)
            if st != 0:
                raise RuntimeError(f"ff_exchange_acquire failed ({st})")
            eng.launches += 1

    def publish(self, blk: RangeBlock) -> None:
        """After a block was filled by hand (peer transport): tell the peers it is complete."""
        if blk.send is None:
            eng = self.engine
            st = eng._lib.ff_exchange_publish(self._xchg, torch.cuda.current_stream(eng.device).cuda_stream)
            if st != 0:
                raise RuntimeError(f"ff_exchange_publish failed ({st})")
            eng.launches += 1

    def finish_arrays(self, pos_local: torch.Tensor, first_exit_local: torch.Tensor, total: int,
                      merge: Optional[MergeFn] = None, counts_local: Optional[torch.Tensor] = None) -> GatheredRange:
        """Same step for results that were not written in place (they came from somewhere else):
        copies them into a block first."""
        blk = self.begin(total, pos_local.device)
        self.acquire(blk)
        n = pos_local.numel()
        blk.pos[:n].copy_(pos_local)
        if counts_local is not None:
            blk.counts[:n].copy_(counts_local)
        blk.first_exit.copy_(first_exit_local.reshape(1))
        self.publish(blk)
        return self.finish(blk, merge, want_counts=counts_local is not None)

    def check(self) -> None:
        """Peer transport: raise if a peer failed to arrive in the last steps (synchronises)."""
        if self._xchg is None:
            return
        eng = self.engine
        status = C.c_int32(0)
        with torch.cuda.device(eng.device):
            st = eng._lib.ff_exchange_status(self._xchg, C.byref(status),
                                             torch.cuda.current_stream(eng.device).cuda_stream)
        if st != 0 or status.value != 0:
            who = (f"rank {status.value - 1} never published its block" if status.value < 0x100
                   else f"rank {status.value - 0x100} never acknowledged an earlier block")
            raise RuntimeError(f"peer-memory exchange: {who} (status {st}); the results of that step are invalid")
