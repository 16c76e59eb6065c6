"""Slab partition plumbing: one process per GPU, torch.distributed only moves the 64-byte IPC handles.

The data path has no NCCL call: axis-0 derivatives read / write the peers' arenas directly over NVLink
(csrc/symm.cu, csrc/elliptic_slab.cu).  What lives here is host logic: the plane / Vec ranges of each
rank (the same arithmetic as GridDesc::init_slab), the handle exchange, and scatter / gather helpers
for tests and benchmarks.
"""
import numpy as np


def slab_range(dim, rank, nranks):
    """(i0, nloc, goff, g_local) of `rank`: planes [i0, i0+nloc) of axis 0 and the local range
    [goff, goff + g_local) of the global Vec (interior nodes, lexicographic; elliptic.C:408-409)."""
    dim = [int(v) for v in dim]
    if dim[0] % nranks:
        raise ValueError("the outermost extent must be divisible by the number of ranks")
    nloc = dim[0] // nranks
    i0 = rank * nloc
    ist0 = int(np.prod([p - 2 for p in dim[1:]])) if len(dim) > 1 else 1
    lo, hi = max(i0, 1), min(i0 + nloc, dim[0] - 1)
    return i0, nloc, (lo - 1) * ist0, max(hi - lo, 0) * ist0


def dirichlet_range(dim, rank, nranks):
    """(doff, nd_local): the local range of the Dirichlet value vector (boundary nodes in walk order)."""
    i0, nloc, goff, gl = slab_range(dim, rank, nranks)
    plane = int(np.prod(dim[1:])) if len(dim) > 1 else 1
    return i0 * plane - goff, nloc * plane - gl


def split_global(vec, dim, nranks, ncomp=1):
    """Cut a global Vec (ncomp values per interior node) into the ranks' local parts."""
    out = []
    for r in range(nranks):
        _, _, goff, gl = slab_range(dim, r, nranks)
        out.append(vec[goff * ncomp:(goff + gl) * ncomp])
    return out


def split_dirichlet(vec, dim, nranks, ncomp=1):
    out = []
    for r in range(nranks):
        doff, nd = dirichlet_range(dim, r, nranks)
        out.append(vec[doff * ncomp:(doff + nd) * ncomp])
    return out


def attach_in_process(ctxs):
    """Several ranks driven by ONE process on one device (tests): map the arenas by plain pointers."""
    for a in ctxs:
        for q, b in enumerate(ctxs):
            if a is not b:
                a.attach_local(q, b)


def exchange_handles(handle, group=None):
    """All-gather one fixed-size bytes object per rank (the 64-byte CUDA IPC handles); works on gloo and nccl."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda()
    handles = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(handles, mine, group=group)
    return [bytes(h.cpu().numpy().tobytes()) for h in handles]


def attach_peers(ctx, group=None):
    """One process per GPU: all-gather the CUDA IPC handles and map every peer's arena."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world != ctx.nranks or rank != ctx.rank:
        raise ValueError("context partition does not match the process group")
    for q, h in enumerate(exchange_handles(ctx.ipc_export(), group)):
        if q != rank:
            ctx.ipc_attach(q, h)
    dist.barrier(group)


def gather_global(local, group=None):
    """Reassemble the global Vec from the ranks' local parts (test / benchmark helper, not on the data path)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    nmax = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros(nmax, dtype=local.dtype, device=local.device)
    buf[:local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)])
