"""Output-neuron sharding of one layer across the ranks of a process group (one process per
GPU).  GPFQ neurons are independent given the layer inputs (rows of U and Q in
step_algorithm.py:141-148 never interact), so each rank solves a contiguous slice of rows and
the slices are exchanged with a single all-gather per layer; the per-neuron squared error norms
ride in the same buffer, so no other collective is needed."""
import torch
import torch.distributed as dist


def _world(group):
    if group is False or not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


world_and_rank = _world


def rows_per_rank(N, world):
    return (N + world - 1) // world


def neuron_slice(N, groups=1, group=None, world=None, rank=None):
    """Contiguous [n0, n1) of the N output neurons owned by this rank (empty for trailing ranks when there are
    fewer units than ranks).  For a grouped convolution the unit is a whole conv group (N/groups neurons): the
    batched grouped solver needs whole groups, and SURVEY.md section 8e partitions grouped layers that way."""
    if world is None:
        world, rank = _world(group)
    unit = N // groups if groups > 1 and N % groups == 0 else 1
    units = N // unit
    per = rows_per_rank(units, world) * unit
    n0 = min(rank * per, N)
    return n0, min(n0 + per, N)


def slice_rows(N, groups, world):
    """Rows per rank of ``neuron_slice`` (the all-gather's per-rank buffer height)."""
    unit = N // groups if groups > 1 and N % groups == 0 else 1
    return rows_per_rank(N // unit, world) * unit


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


def gather_layer(Q, err2, ref2, n0, n1, groups=1, group=None, alphabet=None):
    """All-gather the solved slices so that every rank holds the full Q (N x d) and the full
    per-neuron squared norms.  The layer's only collective.

    ``alphabet`` = (delta 1-element device tensor, K, mode, lam): on CUDA the slice then travels as int8 level
    indices plus the two fp64 norms per neuron (a quarter of the fp32 volume; Q = level * delta exactly), packed by
    ONE kernel and unpacked by ONE kernel (gpfq_pack_slice_f32 / gpfq_unpack_slices_f32).  Without it (CPU tensors
    of the gloo tests, alphabets beyond int8) fp32 rows are exchanged with torch ops."""
    world, _ = _world(group)
    if world == 1:
        return Q, err2, ref2
    N, d = Q.shape
    per = slice_rows(N, groups, world)
    if alphabet is not None and Q.is_cuda:
        from . import _lib
        delta, K, mode, lam = alphabet
        if K + (1 if mode == _lib.MODE_HARD else 0) <= 127:
            rb = int(_lib.lib.gpfq_slice_row_bytes(d))
            mine = torch.empty(per * rb, dtype=torch.uint8, device=Q.device)
            bad = torch.empty(1, dtype=torch.int32, device=Q.device)
            _lib.launch(_lib.lib.gpfq_pack_slice_f32, Q, Q.stride(0), d, n0, n1, per, delta, int(K), int(mode), float(lam),
                        err2, ref2, mine, bad)
            full = torch.empty(world * per * rb, dtype=torch.uint8, device=Q.device)
            dist.all_gather_into_tensor(full, mine, group=group)
            Qf = torch.empty((N, d), dtype=torch.float32, device=Q.device)
            e2 = torch.empty(N, dtype=torch.float64, device=Q.device)
            r2 = torch.empty(N, dtype=torch.float64, device=Q.device)
            _lib.launch(_lib.lib.gpfq_unpack_slices_f32, full, N, d, delta, int(K), int(mode), float(lam), Qf, d, e2, r2)
            return Qf, e2, r2
    mine = pack_slice(Q, err2, ref2, n0, n1, per)
    full = torch.empty((world * per, d + 4), dtype=torch.float32, device=Q.device)
    dist.all_gather_into_tensor(full, mine, group=group)
    return unpack_all(full, N, d)


def gather_inputs(X_local, Xq_local, group=None):
    """Sharded calibration forward: every rank holds the layer inputs of its own images, i.e. a
    contiguous block of calibration rows (rows of the reference's (m x d) matrices, columns of the
    feature-major buffers).  One all-gather of the packed pair gives every rank the full X and X~,
    returned as (m x d) transposed views of feature-major (d x ld) buffers."""
    world, _ = _world(group)
    m_local, d = X_local.shape
    pair = torch.empty((2, d, m_local), dtype=torch.float32, device=X_local.device)
    pair[0].copy_(X_local.t())
    pair[1].copy_(Xq_local.t())
    full = torch.empty((world * 2, d, m_local), dtype=torch.float32, device=X_local.device)
    dist.all_gather_into_tensor(full, pair, group=group)
    full = full.view(world, 2, d, m_local)
    m = world * m_local
    ld = (m + 3) // 4 * 4
    out = torch.zeros((2, d, ld), dtype=torch.float32, device=X_local.device) if ld != m else \
        torch.empty((2, d, ld), dtype=torch.float32, device=X_local.device)
    out[:, :, :m].view(2, d, world, m_local).copy_(full.permute(1, 2, 0, 3))
    return out[0, :, :m].t(), out[1, :, :m].t()
