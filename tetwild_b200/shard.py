"""Multi-GPU plumbing of the hot path: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in
the CPU tests), no collective inside any kernel.

Every query of the path (sample point, candidate face, tet, one-ring, centroid) is independent and only reads the
surface, so the work shards by contiguous index range:
  * the surface (vertices + facets) is broadcast once from rank 0 and each rank builds its own replica of the
    device structure (twg_surface_create / twg_winding_create are deterministic: same input, same tree);
  * rank r owns queries [r*n/G, (r+1)*n/G);
  * the 1-byte decisions (or 8-byte values) are all-gathered; shards differ by at most one element, the tail is padded.
The reference has no counterpart (single process, no MPI/NCCL: SURVEY.md 2.2); this module is the only place where
torch.distributed is used besides bench.py.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """[b, e) of rank `rank` among `world` ranks: contiguous, disjoint, covering [0, n), sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_sizes(n, world):
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


def _dist():
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group)")
    return dist


def broadcast_mesh(V, F, src=0, device=None):
    """Rank `src` passes (V float64 [nV,3], F uint32 [nF,3]); every rank returns identical numpy copies."""
    import torch
    dist = _dist()
    rank = dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    hdr = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == src:
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, 3)
        F = np.ascontiguousarray(F, dtype=np.uint32).reshape(-1, 3)
        hdr[0], hdr[1] = len(V), len(F)
    dist.broadcast(hdr, src)
    nV, nF = int(hdr[0].item()), int(hdr[1].item())
    tv = torch.from_numpy(V.copy()).to(dev) if rank == src else torch.empty((nV, 3), dtype=torch.float64, device=dev)
    tf = torch.from_numpy(F.astype(np.int64)).to(dev) if rank == src else torch.empty((nF, 3), dtype=torch.int64, device=dev)
    dist.broadcast(tv, src)
    dist.broadcast(tf, src)
    return tv.cpu().numpy(), tf.cpu().numpy().astype(np.uint32)


def all_gather_ragged(local, n_total):
    """local: this rank's slice (torch tensor, leading dim = its shard size) of an array of n_total rows split by
    shard_bounds. Returns the full array on every rank (one all_gather of equal-sized, tail-padded pieces)."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, its shard of %d has %d" % (rank, local.shape[0], n_total, sizes[rank]))
    m = max(sizes)
    if m == 0:
        return local.new_empty((0,) + tuple(local.shape[1:]))
    piece = local
    if local.shape[0] < m:
        piece = torch.cat([local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, piece.contiguous())
    if all(s == m for s in sizes):
        return out
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)], 0)


def max_over_ranks(x, device=None):
    """Device-timed milliseconds -> max over ranks (the contract's timing rule)."""
    import torch
    dist = _dist()
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_decisions(n_total, compute_local, device=None):
    """Run `compute_local(b, e) -> uint8 torch tensor of e-b decisions` on this rank's range and return all n_total
    decisions on every rank."""
    dist = _dist()
    b, e = shard_bounds(n_total, dist.get_world_size(), dist.get_rank())
    return all_gather_ragged(compute_local(b, e), n_total)
