"""
Multi-GPU plumbing: one process per GPU, 1-D slab decomposition along the last
spatial axis, periodic ring of ranks.

torch.distributed is used only for the rendezvous (broadcasting the NCCL
unique id) and for host-side collectives in tests; the data-path exchange
(ncclSend/ncclRecv of two contiguous ghost planes per side, ncclAllReduce of
dot products / norms / CFL maxima) lives in libksfd_b200.so.

The pure-index helpers below describe WHICH global planes fill a rank's ghost
planes; they are what the CPU (gloo) tests check bit-exactly against
np.pad(mode='wrap') — the semantics of the reference's DMDA globalToLocal
(KSFD/ksfdsym.py:703-705, 919-920).
"""
import ctypes as C
import glob
import os

import numpy as np

from . import _lib
from .core import SW, dmda_ownership


def nccl_library_path():
    """torch's bundled libnccl.so.2 (same one torch.distributed uses)."""
    try:
        import nvidia.nccl
        base = os.path.dirname(nvidia.nccl.__file__) if getattr(
            nvidia.nccl, '__file__', None) else list(nvidia.nccl.__path__)[0]
        hits = glob.glob(os.path.join(base, 'lib', 'libnccl.so*'))
        if hits:
            return sorted(hits)[0]
    except Exception:
        pass
    return 'libnccl.so.2'


def neighbours(rank, nranks):
    """(down, up) ranks of the periodic slab ring."""
    return (rank - 1) % nranks, (rank + 1) % nranks


def ghost_sources(n_last, nranks, rank, sw=SW):
    """
    Global plane indices that fill this rank's ghost planes:
    returns (lo, hi) — lo = planes start-sw .. start-1, hi = end .. end+sw-1,
    wrapped periodically.  Equal to the indices np.pad(mode='wrap') reads.
    """
    start, count = dmda_ownership(n_last, nranks)[rank]
    lo = [(start - sw + i) % n_last for i in range(sw)]
    hi = [(start + count + i) % n_last for i in range(sw)]
    return lo, hi


def exchange_plan(n_last, nranks, rank, sw=SW):
    """
    The four messages of one halo exchange, as the C library posts them
    (ksfd.cu `exchange`): list of (kind, peer, local plane range / ghost side).
    """
    dn, up = neighbours(rank, nranks)
    start, count = dmda_ownership(n_last, nranks)[rank]
    return [('send', up, (count - sw, count)),     # my top planes -> up's lo
            ('recv', dn, 'lo'),
            ('send', dn, (0, sw)),                 # my bottom planes -> dn's hi
            ('recv', up, 'hi')]


def host_halo_exchange(local, n_last, sw=SW, group=None):
    """
    Reference (host, torch.distributed) version of the slab halo exchange for
    tests: `local` is this rank's slab as a numpy array whose LAST axis is the
    decomposed one.  Returns (lo, hi) ghost slabs.  Works with gloo.
    """
    import torch
    import torch.distributed as dist
    rank, nranks = dist.get_rank(group), dist.get_world_size(group)
    dn, up = neighbours(rank, nranks)
    a = np.ascontiguousarray(np.moveaxis(local, -1, 0))      # planes first
    top = torch.from_numpy(a[-sw:].copy())
    bot = torch.from_numpy(a[:sw].copy())
    lo = torch.empty_like(top)
    hi = torch.empty_like(bot)
    if nranks == 1:
        lo.copy_(top)
        hi.copy_(bot)
    else:
        reqs = [dist.isend(top, up, group=group), dist.irecv(lo, dn, group=group)]
        for r in reqs:
            r.wait()
        reqs = [dist.isend(bot, dn, group=group), dist.irecv(hi, up, group=group)]
        for r in reqs:
            r.wait()
    return (np.moveaxis(lo.numpy(), 0, -1), np.moveaxis(hi.numpy(), 0, -1))


def init_comm(ctx):
    """Create the library's NCCL communicator; rendezvous through the default
    torch.distributed process group (any backend)."""
    import torch
    import torch.distributed as dist
    if ctx.nranks == 1:
        return
    path = nccl_library_path()
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    if ctx.rank == 0:
        _lib.check(lib.ksfd_nccl_unique_id(path.encode(), buf))
    backend = dist.get_backend()
    dev = ctx.tdev if backend == 'nccl' else torch.device('cpu')
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    uid = bytes(t.cpu().tolist())
    ctx.comm_init(path, uid)
    if backend == 'nccl' and os.environ.get('KSFD_HALO_P2P', '1') != '0':
        init_p2p(ctx)


def init_p2p(ctx):
    """Direct NVLink halo push: exchange the CUDA IPC handles of the ranks'
    halo buffers and open the two neighbours'.  Falls back to ncclSend/Recv
    (silently, all ranks together) if any rank cannot export/open a handle."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    buf = C.create_string_buffer(64)
    ok = lib.ksfd_p2p_export(ctx.h, buf) == 0
    mine = torch.tensor(list(buf.raw) + [1 if ok else 0], dtype=torch.uint8, device=ctx.tdev)
    allh = [torch.empty_like(mine) for _ in range(ctx.nranks)]
    dist.all_gather(allh, mine)
    rows = [bytes(t.cpu().tolist()) for t in allh]
    good = all(r[64] == 1 for r in rows)
    good = good and ctx.nranks <= 16
    if good:
        blob = b''.join(r[:64] for r in rows)
        good = lib.ksfd_p2p_import(ctx.h, blob, ctx.nranks) == 0
    flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=ctx.tdev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        ctx.set_option('halo_p2p', 0)
    ctx.halo_p2p = bool(int(flag.item()))
