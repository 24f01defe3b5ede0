"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch).

The path shards by WINDOW (window path) or by VIDEO (frame path) with no data-path collective
(SURVEY.md section 8e); the only exchange step is ONE sum all-reduce of the flat gradient buffer per
step, followed by the fused Adam kernel that applies the 1/world_size scale.  BatchNorm statistics stay
per rank (DDP default), so results parity is asserted at world_size 1 and per rank above it.

The exchange itself: :class:`PeerAllReduce` -- ONE hand-written kernel per rank over NVLink peer memory
(csrc/peer_exchange.cu: reduce-scatter + all-gather by direct peer loads / stores, CUDA IPC mappings set up once
through ``torch.distributed``) -- when all ranks sit on one box and the mappings can be made; the NCCL all-reduce
otherwise (several nodes, IPC refused, ``B200MED_PEER_EXCHANGE=0``).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> tuple:
    """(rank, local_rank, world_size) from the torchrun environment; initialises the process group when
    WORLD_SIZE > 1 (nccl on CUDA, gloo otherwise)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_batch(idx: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Contiguous share of a global batch for one rank.  Even split (sizes differ by at most one), so that a rank's share is
    empty only if the batch has fewer samples than there are ranks -- the loaders drop such a batch on EVERY rank
    (every step ends in a gradient all-reduce: a rank that skipped it would leave the others waiting)."""
    if world_size <= 1:
        return idx
    n = idx.numel()
    return idx[rank * n // world_size:(rank + 1) * n // world_size]


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def max_over_ranks(value: float, device=None) -> float:
    """Step time of the job = slowest rank."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class _DevMem:
    """A raw device allocation as something ``torch.as_tensor`` can alias (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


class PeerAllReduce:
    """In-place sum over all ranks of a flat fp32 buffer of ``n`` elements that lives in peer-mapped memory.

    ``buffer`` is this rank's allocation as a torch tensor (gradients are written straight into it); ``all_reduce()`` launches
    the exchange kernel on the current stream (capturable in a CUDA graph).  Construction is COLLECTIVE (handles travel through
    ``all_gather_object``); it raises on every rank if any rank could not map its peers, so the callers fall back together."""

    def __init__(self, n: int, device):
        import ctypes as C
        from . import _lib
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            raise RuntimeError("PeerAllReduce needs an initialised process group with more than one rank")
        self.rank, self.world, self.n = dist.get_rank(), dist.get_world_size(), int(n)
        self.device = torch.device(device)
        lib = _lib.load()
        self._lib = lib
        self._own, self._imported = [], []
        ok, err = 1, ""
        try:
            if self.world > 16:
                raise RuntimeError("at most 16 ranks")
            buf, flg = C.c_void_p(), C.c_void_p()
            _lib.call("b200med_peer_alloc", self.n * 4, C.byref(buf))
            self._own.append(buf.value)
            _lib.call("b200med_peer_alloc", int(lib.b200med_peer_flag_bytes()), C.byref(flg))
            self._own.append(flg.value)
            hb, hf = C.create_string_buffer(64), C.create_string_buffer(64)
            _lib.call("b200med_peer_export", buf, hb)
            _lib.call("b200med_peer_export", flg, hf)
            mine = (bytes(hb.raw), bytes(hf.raw), os.uname().nodename)
        except Exception as e:      # noqa: BLE001 -- reported collectively below
            ok, err, mine = 0, f"{type(e).__name__}: {e}", (b"", b"", "")
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        bufs, flags = [0] * self.world, [0] * self.world
        if ok and any(e[2] != mine[2] or not e[0] for e in everyone):
            ok, err = 0, "ranks on several hosts (or a rank without a buffer): peer memory is a one-box transport"
        if ok:
            try:
                for q, (hb_q, hf_q, _) in enumerate(everyone):
                    if q == self.rank:
                        bufs[q], flags[q] = self._own[0], self._own[1]
                        continue
                    pb, pf = C.c_void_p(), C.c_void_p()
                    _lib.call("b200med_peer_import", C.create_string_buffer(hb_q, 64), C.byref(pb))
                    self._imported.append(pb.value)
                    _lib.call("b200med_peer_import", C.create_string_buffer(hf_q, 64), C.byref(pf))
                    self._imported.append(pf.value)
                    bufs[q], flags[q] = pb.value, pf.value
            except Exception as e:      # noqa: BLE001
                ok, err = 0, f"{type(e).__name__}: {e}"
        agree = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) == 0:
            self.close()
            raise RuntimeError(f"peer-memory exchange unavailable on rank {self.rank}: {err or 'another rank failed'}")
        self.buffer = torch.as_tensor(_DevMem(self._own[0], self.n), device=self.device)
        self._bufs_dev = torch.tensor(bufs, dtype=torch.int64, device=self.device)
        self._flags_dev = torch.tensor(flags, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier()

    def all_reduce(self):
        import ctypes as C
        from . import _lib
        _lib.call("b200med_peer_allreduce_f32", C.c_void_p(self._bufs_dev.data_ptr()), C.c_void_p(self._flags_dev.data_ptr()),
                  self.rank, self.world, self.n, C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def close(self):
        import ctypes as C
        for p in self._imported:
            self._lib.b200med_peer_close(C.c_void_p(p))
        self._imported = []
        # the own allocations stay alive: tensors alias them (they are a few MB and live as long as the process trains)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
