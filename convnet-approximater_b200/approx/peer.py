"""Peer-memory plumbing for the fused gradient exchange (csrc/peer.cu): one cudaMalloc'ed "symmetric" buffer per rank,
CUDA-IPC handles exchanged once through torch.distributed, every rank maps every peer's buffer (NVLink / NVSwitch P2P).

torch.distributed is used for the 64-byte handle exchange only; the data path is this library's own kernel reading and
writing the mapped pointers.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from approx import backend as B


class _RawView:
    """__cuda_array_interface__ carrier: lets torch wrap a raw device pointer without copying."""

    def __init__(self, ptr, nbytes, typestr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


class SymmetricBuffer:
    """`nbytes` of device memory on this rank plus mapped pointers to the same-sized buffer of every other rank."""

    def __init__(self, nbytes: int, device, group=None):
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        lib = B.lib()
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            B._check(lib.cna_peer_alloc(self.nbytes, ctypes.cast(ctypes.byref(ptr), ctypes.c_void_p), ctypes.cast(handle, ctypes.c_void_p)))
        self.local = ptr.value
        self.peers = [None] * self.world
        self.peers[self.rank] = self.local
        self._opened = []
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            with torch.cuda.device(self.device):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                    out = ctypes.c_void_p()
                    B._check(lib.cna_peer_open(ctypes.cast(buf, ctypes.c_void_p), ctypes.cast(ctypes.byref(out), ctypes.c_void_p)))
                    self.peers[r] = out.value
                    self._opened.append(out.value)
            dist.barrier(group=group)                 # every rank has mapped every buffer before anyone launches on them

    def view(self, byte_offset: int, numel: int, dtype=torch.float32) -> torch.Tensor:
        """Zero-copy torch tensor over [byte_offset, byte_offset + numel * itemsize) of the LOCAL buffer."""
        itemsize = torch.empty((), dtype=dtype).element_size()
        assert byte_offset % itemsize == 0 and byte_offset + numel * itemsize <= self.nbytes
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        with torch.cuda.device(self.device):
            t = torch.as_tensor(_RawView(self.local + byte_offset, numel * itemsize, typestr, (numel,)), device=self.device)
        t._cna_keepalive = self                       # the tensor must not outlive the allocation
        return t

    def peer_array(self):
        return (ctypes.c_void_p * self.world)(*self.peers)

    def close(self):
        lib = B.lib()
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.cna_peer_close(p)
            self._opened = []
            if self.local:
                if self.world > 1:
                    dist.barrier(group=self.group)    # nobody still maps / uses this buffer
                lib.cna_peer_free(self.local)
                self.local = None

    def __del__(self):
        try:
            if self.local and self.world == 1:
                B.lib().cna_peer_free(self.local)
                self.local = None
        except Exception:  # noqa: BLE001
            pass
