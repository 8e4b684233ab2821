"""Peer-memory plumbing for the NVLink data path (csrc/comm.cu): slabs allocated by the library (cudaMalloc) so they can be
exported with CUDA IPC, exposed to torch as zero-copy tensors, and the table of peer-mapped base pointers of every rank."""
import ctypes as C

import torch

from ._lib import lib, check


class Slab:
    """One cudaMalloc'ed allocation viewed as a bf16 torch tensor (zero copy, via __cuda_array_interface__)."""

    def __init__(self, numel, device):
        self.device = torch.device(device)
        self.numel = int(numel)
        self.nbytes = self.numel * 2
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().b200_mem_alloc(C.byref(p), self.nbytes), "b200_mem_alloc")
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": (self.numel,), "typestr": "<i2", "data": (self.ptr, False), "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=self.device).view(torch.bfloat16)
        assert self.tensor.data_ptr() == self.ptr
        self.tensor.zero_()

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib().b200_ipc_export(self.ptr, buf), "b200_ipc_export")
        return buf.raw

    def free(self):
        if self.ptr:
            self.tensor = None
            lib().b200_mem_free(self.ptr)
            self.ptr = 0


class PeerTable:
    """base pointer of `slab` on every rank of the process group, as seen from this rank (own rank: the local pointer)."""

    def __init__(self, slab: Slab, pg):
        import torch.distributed as dist
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        handles = [None] * self.world
        dist.all_gather_object(handles, slab.export_handle(), group=pg)
        self.base = []
        self._opened = []
        with torch.cuda.device(slab.device):
            for j, h in enumerate(handles):
                if j == self.rank:
                    self.base.append(slab.ptr)
                    continue
                p = C.c_void_p()
                check(lib().b200_ipc_import(h, C.byref(p)), f"b200_ipc_import(rank {j})")
                self.base.append(p.value)
                self._opened.append(p.value)

    def close(self):
        for p in self._opened:
            lib().b200_ipc_close(p)
        self._opened = []
