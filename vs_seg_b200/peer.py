"""Peer-memory accumulator of the multi-GPU sliding window (one node, one process per GPU).

SURVEY.md §8e: windows are independent, so rank r runs a contiguous block of the MONAI window list.  Instead of
blending into a private accumulator and reducing 189 MB per volume afterwards, every rank blends its windows
STRAIGHT into the destination rank's accumulator over NVLink: the last kernel of the network
(vsseg_conv3d_gate_logits) issues red.global.add.f32 on a buffer that the destination allocated and the others
mapped with CUDA IPC.  The exchange therefore overlaps the compute tile by tile and there is no reduce pass, no
per-rank accumulator and no per-rank zero fill.

Hand-shake per volume k (buffer j = k % 2, all flags live in the destination's memory, system-scope acquire/release):
  every rank   wait  release[j] >= k-1      (volume k-2, the previous user of buffer j, was finalised and re-zeroed)
               blend its windows into acc[j] (captured graph)
               set   arrive[j][rank] = k+1
  destination  wait  arrive[j][*] >= k+1  ->  finalise acc[j] (divide, argmax mask, Dice sums)  ->  zero acc[j]
               set   release[j] = k+1
Two buffers let the ranks run one volume ahead of the destination's finalise.  The waits are bounded spins
(vsseg_flag_wait): a dead peer raises instead of hanging the GPU.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.distributed as dist

from . import lib as _lib

_TIMEOUT_CYCLES = int(20e9)   # ~10 s of SM clocks


class _DevMem:
    """__cuda_array_interface__ holder: lets torch alias memory this module allocated (destination rank only)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _alias(ptr, shape, dtype, device):
    typestr = {torch.float32: "<f4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_DevMem(ptr, shape, typestr), device=device)


class PeerUnavailable(RuntimeError):
    """CUDA IPC mapping of the destination's accumulator failed on some rank (every rank raises together)."""


class PeerAccumulator:
    def __init__(self, shape, device, group=None, dst=0):
        self.lib = _lib.load()
        self.shape = tuple(int(v) for v in shape)
        self.device = torch.device(device)
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.k = 0
        nbytes = 4 * math.prod(self.shape)
        nflags = 2 * self.world + 2
        sizes = [nbytes, nbytes, 8 * max(nflags, 16)]
        self._owned, self._opened = [], []
        handles, ok = None, True
        if self.rank == dst:
            handles = []
            try:
                for sz in sizes:
                    p, h = C.c_void_p(), C.create_string_buffer(64)
                    _lib.check(self.lib.vsseg_peer_alloc(sz, C.byref(p), h), "peer_alloc")
                    self._owned.append(p.value)
                    handles.append(h.raw)
            except _lib.NativeLibraryError:
                handles, ok = None, False
        box = [handles]
        dist.broadcast_object_list(box, src=dst, group=group)
        ptrs = list(self._owned)
        if box[0] is None:
            ok = False
        elif self.rank != dst:
            ptrs = []
            try:
                for h in box[0]:
                    p = C.c_void_p()
                    _lib.check(self.lib.vsseg_peer_open(C.create_string_buffer(h, 64), C.byref(p)), "peer_open")
                    self._opened.append(p.value)
                    ptrs.append(p.value)
            except _lib.NativeLibraryError:
                ok = False
        # every rank must agree: one failed mapping sends all of them to the reduce path
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            raise PeerUnavailable("CUDA IPC mapping of the peer accumulator is not available on this node")
        self.acc_ptr = ptrs[:2]
        self.flags_ptr = ptrs[2]
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.acc = [_alias(p, self.shape, torch.float32, self.device) for p in self.acc_ptr] if self.rank == dst else None
        dist.barrier(group=group)   # every rank has mapped the buffers before anyone blends

    # flag addresses (int64 each): arrive[j][r] at j*world + r, release[j] at 2*world + j
    def _arrive(self, j, r=0):
        return self.flags_ptr + 8 * (j * self.world + r)

    def _release(self, j):
        return self.flags_ptr + 8 * (2 * self.world + j)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def begin(self):
        """Start volume k on this rank: returns the base address of the accumulator to blend into."""
        j = self.k % 2
        if self.k >= 2:
            _lib.check(self.lib.vsseg_flag_wait(self._release(j), 1, self.k - 1, _TIMEOUT_CYCLES, self.err.data_ptr(),
                                                self._stream()), "flag_wait")
        return self.acc_ptr[j]

    def arrive(self):
        """This rank's blends of volume k are queued: publish them."""
        j = self.k % 2
        _lib.check(self.lib.vsseg_flag_set(self._arrive(j, self.rank), self.k + 1, self._stream()), "flag_set")

    def gather(self):
        """Destination rank: wait for every rank's blends of volume k; returns the accumulator tensor."""
        j = self.k % 2
        _lib.check(self.lib.vsseg_flag_wait(self._arrive(j), self.world, self.k + 1, _TIMEOUT_CYCLES, self.err.data_ptr(),
                                            self._stream()), "flag_wait")
        return self.acc[j]

    def release(self):
        """Destination rank: the accumulator of volume k has been consumed: re-zero it and hand it back."""
        j = self.k % 2
        self.acc[j].zero_()
        _lib.check(self.lib.vsseg_flag_set(self._release(j), self.k + 1, self._stream()), "flag_set")

    def end(self):
        self.k += 1

    def check(self):
        """Host check of the bounded waits (synchronises; call outside the hot loop)."""
        if int(self.err.item()):
            raise _lib.NativeLibraryError("peer accumulator: a rank did not arrive within the timeout")

    def close(self):
        for p in self._opened:
            self.lib.vsseg_peer_close(p)
        for p in self._owned:
            self.lib.vsseg_peer_free(p)
        self._opened, self._owned = [], []
