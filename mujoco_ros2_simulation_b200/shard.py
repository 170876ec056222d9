"""Sharding of worlds / sensor-pose batches over the GPUs of one box, and the gather of the
results over NVLink (BASELINE.json configs[4]).

[NOT IN REFERENCE] -- the reference is one process on one CPU.  One process per GPU
(``torchrun``); worlds are independent, so nothing is exchanged while casting.  The only
cross-GPU traffic is the result gather, in one of two forms:

``PeerGather``  (the product path)  the gathered arrays live in PEER BUFFERS on the root GPU
                (``b2_peer_buffer_*``, CUDA IPC) that every rank maps.  Two ways to fill them:
                ``ptr()`` -- hand the cast kernels a pointer into the root's memory, so the
                kernels' own float4 / int4 result stores ARE the gather; ``push()`` -- cast into
                local memory block by block and let the copy engine carry each finished block
                across NVLink while the next one is cast (bulk transfers: the better form when
                the root's ingress is the limit, i.e. at 8 GPUs).
``BlockGather`` (the library baseline, and the CPU test tier over gloo)  the rank's scans are cut
                into blocks; block k is all-gathered with ``torch.distributed`` on a side
                stream while block k+1 is being cast.

PyTorch is plumbing here (process group, streams); the data path is libb2sense.so's kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi


# ------------------------------------------------------------------------------------------------
# who owns what
# ------------------------------------------------------------------------------------------------
def world_block(nworld: int, nrank: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of the ``nworld`` worlds owned by ``rank``: block sizes differ by
    at most one, earlier ranks get the larger blocks, every world has exactly one owner."""
    if nworld < 0 or nrank < 1 or not 0 <= rank < nrank:
        raise ValueError(f"world_block({nworld}, {nrank}, {rank})")
    q, r = divmod(nworld, nrank)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def owner_of(world: int, nworld: int, nrank: int) -> int:
    """Rank that owns ``world`` under ``world_block``."""
    if not 0 <= world < nworld:
        raise ValueError(f"world {world} out of range")
    q, r = divmod(nworld, nrank)
    edge = r * (q + 1)
    return world // (q + 1) if world < edge else r + (world - edge) // max(q, 1)


def scan_blocks(nscan: int, nblock: int) -> list[tuple[int, int]]:
    """Cut a rank's ``nscan`` scans into at most ``nblock`` near-equal contiguous blocks (the unit of
    the cast / gather pipeline)."""
    nblock = max(1, min(nblock, nscan)) if nscan > 0 else 1
    return [(nscan * k // nblock, nscan * (k + 1) // nblock) for k in range(nblock)
            if nscan * (k + 1) // nblock > nscan * k // nblock]


def padded_rows(nworld: int, nrank: int) -> int:
    """Rows every rank contributes to an equal-sized collective: the largest block."""
    return -(-nworld // nrank)


# ------------------------------------------------------------------------------------------------
# BlockGather: torch.distributed all-gather per block, overlapped with the next block's cast
# ------------------------------------------------------------------------------------------------
@dataclass
class GatheredArray:
    """One gathered array in block-major layout [block][rank][rows of the block][row_elems];
    ``assemble()`` restores the global order [rank][row][row_elems]."""
    data: "object"  # torch tensor, flat
    blocks: list
    nrank: int
    row_elems: int

    def assemble(self):
        import torch
        rows = sum(b - a for a, b in self.blocks)
        out = torch.empty((self.nrank, rows, self.row_elems), dtype=self.data.dtype, device=self.data.device)
        off = 0
        for a, b in self.blocks:
            n = (b - a) * self.row_elems
            chunk = self.data[off:off + self.nrank * n].view(self.nrank, b - a, self.row_elems)
            out[:, a:b] = chunk
            off += self.nrank * n
        return out


class BlockGather:
    """All-gather of per-rank result arrays, block by block.

    ``run(produce)`` calls ``produce(k, lo, hi)`` for every block -- it must enqueue, on the
    current stream, the work that fills rows [lo, hi) of the local arrays -- and all-gathers block k
    on ``comm_stream`` while block k+1 is produced.  On CPU tensors (gloo, the test tier) it is
    the same code without streams."""

    def __init__(self, local: dict, nrows: int, nblock: int = 4, group=None, comm_stream=None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.nrank = dist.get_world_size(group)
        self.local = local
        self.blocks = scan_blocks(nrows, nblock)
        self.comm_stream = comm_stream
        self.out = {}
        for name, t in local.items():
            if t.shape[0] != nrows:
                raise ValueError(f"{name}: {t.shape[0]} rows, expected {nrows}")
            row = int(np.prod(t.shape[1:])) if t.dim() > 1 else 1
            flat = torch.empty(self.nrank * nrows * row, dtype=t.dtype, device=t.device)
            self.out[name] = GatheredArray(flat, self.blocks, self.nrank, row)
        self.cuda = any(t.is_cuda for t in local.values())

    def run(self, produce):
        import torch
        off = {name: 0 for name in self.local}
        pending = []
        for k, (lo, hi) in enumerate(self.blocks):
            produce(k, lo, hi)
            ev = None
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
            ctx = torch.cuda.stream(self.comm_stream) if (self.cuda and self.comm_stream is not None) else _Null()
            with ctx:
                if ev is not None and self.comm_stream is not None:
                    self.comm_stream.wait_event(ev)
                for name, t in self.local.items():
                    ga = self.out[name]
                    n = (hi - lo) * ga.row_elems
                    dst = ga.data[off[name]:off[name] + self.nrank * n]
                    src = t[lo:hi].reshape(-1)
                    w = self.dist.all_gather_into_tensor(dst, src, group=self.group, async_op=True)
                    pending.append(w)
                    off[name] += self.nrank * n
        for w in pending:
            w.wait()  # CUDA: orders the current stream after the collective; CPU: blocks
        return self.out


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ------------------------------------------------------------------------------------------------
# PeerGather: the kernels' own stores are the gather (P2P over NVLink)
# ------------------------------------------------------------------------------------------------
class _DevMem:
    """Expose raw device memory to torch / numpy consumers (``torch.as_tensor(_DevMem(...))``)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


_TYPESTR = {"float32": "<f4", "int32": "<i4", "uint16": "<u2", "float64": "<f8", "uint8": "|u1"}


class PeerGather:
    """Gather-to-root fused into the producing kernels.

    ``arrays``: name -> (row_elems, dtype name).  Every rank contributes ``rows`` rows per array.
    The root allocates one peer buffer per array holding [nrank][rows][row_elems]; the other ranks
    map it.  ``ptr(name, row=0)`` is the device pointer (an int, valid in this process) of this
    rank's slice -- pass it as the output argument of the ``_device`` entry points.  ``root_tensor``
    (root only) views the whole gathered array as a torch tensor."""

    def __init__(self, arrays: dict, rows: int, rank: int, nrank: int, device: int, root: int = 0, group=None):
        import torch
        import torch.distributed as dist
        self.lib = _capi.load()
        self.rank, self.nrank, self.root, self.device, self.rows = rank, nrank, root, device, rows
        self.arrays = dict(arrays)
        self.base = {}
        self._own = rank == root
        handles = torch.zeros((len(arrays), _capi.PEER_HANDLE_BYTES), dtype=torch.uint8)
        for i, (name, (row_elems, dt)) in enumerate(self.arrays.items()):
            nbytes = nrank * rows * row_elems * np.dtype(dt).itemsize
            if self._own:
                p = C.c_void_p()
                h = (C.c_uint8 * _capi.PEER_HANDLE_BYTES)()
                _capi.check(self.lib.b2_peer_buffer_create(device, nbytes, C.byref(p), h))
                self.base[name] = int(p.value)
                handles[i] = torch.frombuffer(bytearray(h), dtype=torch.uint8)
        # the handle travels through the process group (device tensor for NCCL, host for gloo)
        backend = dist.get_backend(group)
        hdev = handles.to(torch.device("cuda", device)) if backend == "nccl" else handles
        dist.broadcast(hdev, src=root, group=group)
        handles = hdev.cpu()
        if not self._own:
            for i, name in enumerate(self.arrays):
                p = C.c_void_p()
                h = (C.c_uint8 * _capi.PEER_HANDLE_BYTES).from_buffer_copy(bytes(handles[i].numpy().tobytes()))
                _capi.check(self.lib.b2_peer_buffer_open(device, h, C.byref(p)))
                self.base[name] = int(p.value)

    def row_bytes(self, name: str) -> int:
        row_elems, dt = self.arrays[name]
        return row_elems * np.dtype(dt).itemsize

    def ptr(self, name: str, row: int = 0) -> int:
        """device pointer of row ``row`` of THIS rank's slice of ``name`` (root's memory)"""
        return self.base[name] + (self.rank * self.rows + row) * self.row_bytes(name)

    def push(self, scene, name: str, row: int, src):
        """Bulk form: copy the finished rows `src` (a contiguous device tensor, rows [row, row + n) of
        this rank's slice) into the root's buffer with the copy engine, behind the work enqueued on
        the scene's stream so far (b2_scene_copy_async).  On the root itself it is a local copy."""
        scene.copy_async(self.ptr(name, row), src, src.numel() * src.element_size())

    def root_tensor(self, name: str):
        import torch
        if not self._own:
            raise RuntimeError("root_tensor() on a non-root rank")
        row_elems, dt = self.arrays[name]
        return torch.as_tensor(_DevMem(self.base[name], (self.nrank, self.rows, row_elems), _TYPESTR[dt]),
                               device=torch.device("cuda", self.device))

    def bytes_per_rank(self) -> int:
        return sum(self.rows * self.row_bytes(n) for n in self.arrays)

    def close(self):
        for name, p in list(self.base.items()):
            if self._own:
                self.lib.b2_peer_buffer_destroy(self.device, C.c_void_p(p))
            else:
                self.lib.b2_peer_buffer_close(self.device, C.c_void_p(p))
        self.base = {}
