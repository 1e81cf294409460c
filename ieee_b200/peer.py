"""Exchange buffers of a gallery sharded over the GPUs of one box (include/ieee_b200.h, "Peer exchange").

Every rank of the gallery group owns one buffer (``ieee_peer_alloc``) and maps the buffers of its peers through
cudaIpc handles exchanged once over the process group; the rank kernels then store lists / partial counts / results
straight into peer memory over NVLink.  Buffers are pooled per (group, device) and only ever grow, because an
evaluator is typically rebuilt for every evaluation while the exchange shapes stay the same.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _Raw:
    """Raw device memory as a CUDA-array-interface object (torch.as_tensor wraps it without copying)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerLink:
    """This rank's exchange buffer plus the mapped buffers of the other ranks of `group`."""

    def __init__(self, group, device: torch.device, nbytes: int):
        import torch.distributed as dist
        lib = _lib.load()
        self.group, self.device = group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nbytes = nbytes
        self.epoch = 0
        own, handle = C.c_void_p(0), C.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.check(lib.ieee_peer_alloc(nbytes, C.byref(own), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)      # host-side, once per buffer
            self.base = []
            for s, h in enumerate(handles):
                if s == self.rank:
                    self.base.append(own.value)
                else:
                    p = C.c_void_p(0)
                    _lib.check(lib.ieee_peer_open(C.create_string_buffer(h, 64), C.byref(p)))
                    self.base.append(p.value)
        self.own = own.value
        self.view = torch.as_tensor(_Raw(self.own, nbytes), device=device)          # uint8 view of the own buffer

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def descriptor(self, Qb_max, Qb, Qtot, q_base, cap, W, epoch) -> _lib.PeerExchange:
        ex = _lib.PeerExchange()
        ex.shards, ex.my_shard, ex.epoch = self.world, self.rank, epoch
        for s, b in enumerate(self.base):
            ex.base[s] = b
        ex.Qb_max, ex.Qb, ex.Qtot, ex.q_base, ex.cap, ex.W = Qb_max, Qb, Qtot, q_base, cap, W
        return ex


class LocalPeers:
    """`shards` exchange buffers on ONE device, no IPC: every virtual rank sees all of them (tests drive the protocol
    with one stream per virtual rank)."""

    def __init__(self, shards: int, device: torch.device, nbytes: int):
        lib = _lib.load()
        self.base = []
        with torch.cuda.device(device):
            for _ in range(shards):
                p, handle = C.c_void_p(0), C.create_string_buffer(64)
                _lib.check(lib.ieee_peer_alloc(nbytes, C.byref(p), handle))
                self.base.append(p.value)
        self.shards, self.nbytes, self.device = shards, nbytes, device
        self.views = [torch.as_tensor(_Raw(b, nbytes), device=device) for b in self.base]

    def descriptor(self, my, Qb_max, Qb, Qtot, q_base, cap, W, epoch) -> _lib.PeerExchange:
        ex = _lib.PeerExchange()
        ex.shards, ex.my_shard, ex.epoch = self.shards, my, epoch
        for s, b in enumerate(self.base):
            ex.base[s] = b
        ex.Qb_max, ex.Qb, ex.Qtot, ex.q_base, ex.cap, ex.W = Qb_max, Qb, Qtot, q_base, cap, W
        return ex

    def free(self):
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        self.views = []
        for b in self.base:
            lib.ieee_peer_free(b)
        self.base = []


_POOL = {}


def link_for(group, device: torch.device, nbytes: int) -> PeerLink:
    """The pooled link of (group, device), re-created (a collective: every rank of the group gets here with the same
    size) when a larger buffer is needed.  Old buffers stay mapped: a peer may still be reading them."""
    key = (id(group), str(device))
    link = _POOL.get(key)
    if link is None or link.nbytes < nbytes:
        grown = max(nbytes, 2 * link.nbytes if link is not None else 0)
        new = PeerLink(group, device, grown)
        if link is not None:
            new.epoch = link.epoch
            new._previous = link                       # keep the old mappings alive
        _POOL[key] = link = new
    return link
