"""Output-neuron sharding of one layer across the ranks of a process group (one process per
GPU).  GPFQ neurons are independent given the layer inputs (rows of U and Q in
step_algorithm.py:141-148 never interact), so each rank solves a contiguous slice of rows and
the slices are exchanged with a single all-gather per layer; the per-neuron squared error norms
ride in the same buffer, so no other collective is needed."""
import torch
import torch.distributed as dist


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def rows_per_rank(N, world):
    return (N + world - 1) // world


def neuron_slice(N, groups=1, group=None, world=None, rank=None):
    """Contiguous [n0, n1) of the N output neurons owned by this rank (empty for trailing ranks
    when N < world).  Slices may cut through conv groups; the solver intersects them per group."""
    if world is None:
        world, rank = _world(group)
    per = rows_per_rank(N, world)
    n0 = min(rank * per, N)
    return n0, min(n0 + per, N)


def pack_slice(Q, err2, ref2, n0, n1, per):
    """[per x (d + 4)] fp32: the Q rows of the slice, then ||u_n||^2 and ||X w_n||^2 as raw float64
    bit patterns (two fp32 words each); rows beyond the slice are zero."""
    d = Q.shape[1]
    buf = torch.zeros((per, d + 4), dtype=torch.float32, device=Q.device)
    rows = n1 - n0
    if rows > 0:
        buf[:rows, :d] = Q[n0:n1]
        buf[:rows, d:d + 2] = err2[n0:n1].contiguous().view(torch.float32).view(rows, 2)
        buf[:rows, d + 2:d + 4] = ref2[n0:n1].contiguous().view(torch.float32).view(rows, 2)
    return buf


def unpack_all(full, N, d):
    """Inverse of pack_slice over the concatenation of all ranks' buffers."""
    full = full.view(-1, d + 4)[:N]
    Q = full[:, :d].contiguous()
    err2 = full[:, d:d + 2].contiguous().view(torch.float64).view(N)
    ref2 = full[:, d + 2:d + 4].contiguous().view(torch.float64).view(N)
    return Q, err2, ref2


def gather_layer(Q, err2, ref2, n0, n1, groups=1, group=None):
    """All-gather the solved slices so that every rank holds the full Q (N x d) and the full
    per-neuron squared norms.  The layer's only collective."""
    world, _ = _world(group)
    if world == 1:
        return Q, err2, ref2
    N, d = Q.shape
    per = rows_per_rank(N, world)
    mine = pack_slice(Q, err2, ref2, n0, n1, per)
    full = torch.empty((world * per, d + 4), dtype=torch.float32, device=Q.device)
    dist.all_gather_into_tensor(full, mine, group=group)
    return unpack_all(full, N, d)
