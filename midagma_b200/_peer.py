"""Peer-visible exchange buffers for the persistent kernels that sum partial results over the GPUs of one box by plain
loads / stores over NVLink (csrc/peer.cu, csrc/lin_iter.cu, csrc/mlp_iter.cu): one allocation per GPU, mapped into every
process of the group through CUDA IPC handles.  The reference has no multi-device path; this is the plumbing behind
``DagmaLinear.shard_rows`` / ``DagmaNonlinear.group``."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class PeerExchange:
    """``PeerExchange.create(group, nbytes, device, want)`` returns an instance -- or ``None`` on EVERY rank when any rank
    does not want the peer path (shape not supported, switched off) or any mapping fails: two small all-reduces make
    the decision collective.  ``ptrs[r]`` is rank r's buffer as mapped into this process (zeroed; the second all-reduce
    is also the barrier after which every buffer is mapped everywhere)."""

    def __init__(self, group, rank, world, own, imported, ptrs):
        self.group, self.rank, self.world = group, rank, world
        self._own, self._imported, self.ptrs = own, imported, ptrs

    @staticmethod
    def _agree(group, device, ok: bool) -> bool:
        import torch.distributed as dist
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64,
                         device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        return bool(t.item() > 0.5)

    @classmethod
    def create(cls, group, nbytes: int, device, want: bool):
        import torch.distributed as dist
        lib = _lib.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        want = want and dist.get_backend(group) == "nccl" and 2 <= world <= 8 and nbytes > 0
        if not cls._agree(group, device, want):
            return None
        own, handle = C.c_void_p(), (C.c_ubyte * 64)()
        ok = lib.dagma_peer_alloc(nbytes, C.byref(own)) == 0
        ok = ok and lib.dagma_peer_export(own, handle) == 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle) if ok else None, group=group)
        ptrs, imported = (C.c_void_p * world)(), []
        ok = ok and all(h is not None for h in handles)
        if ok:
            for r in range(world):
                if r == rank:
                    ptrs[r] = own.value
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                if lib.dagma_peer_import(buf, C.byref(p)) != 0:
                    ok = False
                    break
                imported.append(p)
                ptrs[r] = p.value
        if not cls._agree(group, device, ok):
            for p in imported:
                lib.dagma_peer_release(p)
            if own.value:
                lib.dagma_peer_free(own)
            return None
        return cls(group, rank, world, own, imported, ptrs)

    def close(self) -> None:
        """Release the mappings and free the own buffer (collective: every rank of the group calls it)."""
        import torch.distributed as dist
        lib = _lib.load()
        torch.cuda.synchronize()
        for p in self._imported:
            lib.dagma_peer_release(p)
        self._imported = []
        dist.barrier(group=self.group)           # nobody maps the buffer any more
        if self._own is not None:
            lib.dagma_peer_free(self._own)
            self._own = None
