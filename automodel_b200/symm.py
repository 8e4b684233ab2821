"""Symmetric-memory plumbing for the NVLink / NVSwitch data path (csrc/comm.cu): the flat parameter and gradient storage of all units
is allocated with torch's symmetric-memory allocator (cuMemCreate + handle exchange + NVLS multicast binding - plumbing, like the
caching allocator), rendezvoused once over the shard group, and handed to a `b200_ctx` as raw pointers.  All data movement on these
buffers is done by this repository's kernels (b200_reducescatter_layer / b200_allgather_layer)."""
import ctypes as C

import torch

from ._lib import lib, check


class SymmetricSlab:
    """`numel` elements of `dtype`, the same allocation on every rank of `pg`; `.tensor` is the local view, `.peer_ptrs[j]` this rank's
    mapping of rank j's copy, `.multicast_ptr` the NVLS mapping (0 when the platform has no multicast support)."""

    def __init__(self, numel, dtype, device, pg):
        import torch.distributed._symmetric_memory as symm_mem
        self.tensor = symm_mem.empty(int(numel), dtype=dtype, device=device)
        self.tensor.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm_mem.rendezvous(self.tensor, group=pg.group_name if hasattr(pg, "group_name") else pg)
        self.nbytes = self.tensor.numel() * self.tensor.element_size()
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.multicast_ptr = int(self.handle.multicast_ptr) if getattr(self.handle, "has_multicast_support", False) and self.handle.multicast_ptr else 0
        # rendezvous may return a handle whose pointers address the start of the underlying allocation: account for the tensor's offset
        off = int(getattr(self.handle, "offset", 0) or 0)
        self.peer_ptrs = [p + off for p in self.peer_ptrs]
        if self.multicast_ptr:
            self.multicast_ptr += off
        rank = self.handle.rank
        assert self.peer_ptrs[rank] == self.tensor.data_ptr(), (self.peer_ptrs[rank], self.tensor.data_ptr(), off)


class CommContext:
    """b200_ctx of one rank: slot 0 = parameter slab, slot 1 = gradient slab, plus the signal pad."""

    PARAMS, GRADS = 0, 1

    def __init__(self, pg, device):
        import torch.distributed as dist
        self.pg = pg
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        self.device = torch.device(device)
        h = C.c_void_p()
        check(lib().b200_ctx_create(C.byref(h), self.rank, self.world), "b200_ctx_create")
        self.ptr = h
        self._slabs = {}
        pad_bytes = int(lib().b200_ctx_signal_pad_bytes())
        self.pad = SymmetricSlab(pad_bytes // 4, torch.int32, self.device, pg)
        arr = (C.c_void_p * self.world)(*self.pad.peer_ptrs)
        check(lib().b200_ctx_set_signal_pad(self.ptr, arr, pad_bytes), "b200_ctx_set_signal_pad")
        dist.barrier(group=pg)      # every rank's pad is zeroed before anyone's kernel can signal into it

    def register(self, slot, slab: SymmetricSlab):
        arr = (C.c_void_p * self.world)(*slab.peer_ptrs)
        check(lib().b200_ctx_register_buffer(self.ptr, int(slot), arr, C.c_void_p(slab.multicast_ptr or None), slab.nbytes), "b200_ctx_register_buffer")
        self._slabs[slot] = slab

    def has_multicast(self, slot):
        return bool(lib().b200_ctx_has_multicast(self.ptr, int(slot)))

    def set_timeout_ms(self, ms):
        check(lib().b200_ctx_set_timeout_ms(self.ptr, int(ms)), "b200_ctx_set_timeout_ms")

    def close(self):
        if self.ptr:
            lib().b200_ctx_destroy(self.ptr)
            self.ptr = None
