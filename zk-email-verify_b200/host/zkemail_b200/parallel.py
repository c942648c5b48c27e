"""Multi-GPU plumbing: the EmailVerifier path shards by PROOFS - one process per GPU, the proving
key replicated, every rank proves a contiguous slice of the batch, and the only exchange is the final gather of
256-byte proofs + public signals.  No data-path collective exists; torch.distributed carries the gather (NCCL on
GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of [0, n_items) owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_proofs(local_proofs: bytes, local_publics: bytes, n_items: int, n_public: int, group=None):
    """All ranks call this with the proofs of their shard (in shard order); every rank returns the full, ordered
    (proofs, publics) byte strings.  Uses fixed-size all_gather on uint8 tensors (shards padded to the largest)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_range(n_items, rank, world)
    if len(local_proofs) != 256 * len(mine) or len(local_publics) != 32 * n_public * len(mine):
        raise ValueError("local buffers do not match this rank's shard")
    max_items = (n_items + world - 1) // world
    rec = 256 + 32 * n_public
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    buf = torch.zeros(max_items * rec, dtype=torch.uint8, device=device)
    packed = bytearray()
    for i in range(len(mine)):
        packed += local_proofs[256 * i:256 * (i + 1)] + local_publics[32 * n_public * i:32 * n_public * (i + 1)]
    if packed:
        buf[: len(packed)] = torch.frombuffer(packed, dtype=torch.uint8).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    proofs, publics = bytearray(), bytearray()
    for r in range(world):
        data = bytes(out[r].cpu().numpy())
        for i in range(len(shard_range(n_items, r, world))):
            proofs += data[rec * i: rec * i + 256]
            publics += data[rec * i + 256: rec * (i + 1)]
    return bytes(proofs), bytes(publics)


# ---------------------------------------------------------------------------------------------------------------------
# Intra-proof sharding (BASELINE configs[3] / [4]): ONE proof across 2, 4 or 8 GPUs.
# The engine does the per-GPU work (zke_shard_begin / _mid / _end, include/zkemail_b200.h); the exchanges between the
# steps are collectives on the engine's device vectors, carried by torch.distributed (NCCL over NVLink on GPUs):
#   exchange 1 / 2 : all-to-all "columns -> rows" / "rows -> columns" of the three N x 32-byte evaluation vectors -
#                    every GPU sends (world - 1) / world^2 of a vector to the others: 3 x 2 x N x 32 x (world-1)/world
#                    bytes cross NVLink per proof in total (352 MiB at N = 2^24, world = 8)
#   all-gather     : 388 bytes per GPU (five partial points + the constraint-check word)
# The index logic of the exchanges lives in shard_exchange() and works on any tensor (CPU tensors in the gloo tests).
def shard_views(vec, rank: int, world: int):
    """Views of an [N, ...] tensor: (rows, cols) where rows[q] = (block rank, columns of q) and cols[g] = (block g,
    columns of rank) - the pieces GPU `rank` exchanges with GPU q / g."""
    n = vec.shape[0]
    if n % (world * world):
        raise ValueError("vector length must be a multiple of world^2")
    m, c = n // world, n // (world * world)
    rows = [vec[rank * m + q * c: rank * m + (q + 1) * c] for q in range(world)]
    cols = [vec[g * m + rank * c: g * m + (rank + 1) * c] for g in range(world)]
    return rows, cols


def shard_exchange(vec, rank: int, world: int, to_rows: bool, group=None):
    """In-place all-to-all on a vector laid out at global positions: to_rows=True sends this GPU's column range of every
    block g to GPU g (afterwards the GPU holds its whole row block); to_rows=False is the inverse."""
    import torch.distributed as dist
    rows, cols = shard_views(vec, rank, world)
    send, recv = (cols, rows) if to_rows else (rows, cols)
    # the diagonal piece (block rank, columns of rank) is both sent to and received from this GPU itself: clone the
    # send side so that source and destination never alias
    send = [t.clone() for t in send]
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv, send, group=group)       # one grouped NCCL all-to-all over NVLink
        return
    # backends without all-to-all (gloo in the CPU tests): the same pattern as paired sends / receives
    recv[rank].copy_(send[rank])
    ops = []
    for peer in range(world):
        if peer != rank:
            ops.append(dist.P2POp(dist.isend, send[peer], peer, group))
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer, group))
    for work in dist.batch_isend_irecv(ops):
        work.wait()


class _DevVec:
    """Zero-copy torch view of an engine-owned device vector (N x 32 bytes) via __cuda_array_interface__."""

    def __init__(self, ptr: int, n_elems: int):
        self.__cuda_array_interface__ = {"shape": (n_elems, 32), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def prove_sharded(ctx, packed_input: bytes | None, rs: bytes | None = None, group=None):
    """One Groth16 proof computed by all ranks of `group` together (every rank passes the same input and, for a
    reproducible proof, the same 64-byte (r, s)).  packed_input = None: the witness already resident in the context
    (zke_witness / zke_load_witness) is proved.  Returns (proof bytes, public signal bytes, status) on every rank."""
    import ctypes
    import torch
    import torch.distributed as dist
    from . import _lib as L
    from .engine import AssertFailed
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    err = ctypes.create_string_buffer(L.ERRCAP)

    def ok(rc):
        if rc != 0:
            raise L.ZkeError(err.value.decode())

    if packed_input is not None:
        ctx.witness(packed_input, 1, want_witness=False, raise_on_fail=False)   # replicated: every GPU needs the witness
    ok(L.zke_shard_begin(ctx.handle, rank, world, err, L.ERRCAP))
    n = L.c_size_t()
    vecs = []
    for which in range(3):
        ptr = L.zke_shard_vector(ctx.handle, which, ctypes.byref(n))
        vecs.append(torch.as_tensor(_DevVec(ptr, n.value), device=torch.device("cuda", ctx.device)))
    for v in vecs:
        shard_exchange(v, rank, world, True, group)
    torch.cuda.synchronize()
    ok(L.zke_shard_mid(ctx.handle, err, L.ERRCAP))
    for v in vecs:
        shard_exchange(v, rank, world, False, group)
    torch.cuda.synchronize()
    partial = ctypes.create_string_buffer(L.SHARD_PARTIAL_BYTES)
    publics = ctypes.create_string_buffer(max(1, 32 * ctx.n_public))
    ok(L.zke_shard_end(ctx.handle, partial, publics, err, L.ERRCAP))
    mine = torch.frombuffer(bytearray(partial.raw), dtype=torch.uint8).cuda(ctx.device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    allp = b"".join(bytes(t.cpu().numpy()) for t in gathered)
    return combine_partials(ctx.zkey, allp, world, rs, publics.raw[: 32 * ctx.n_public])


def combine_partials(zkey, partials: bytes, world: int, rs: bytes | None, publics: bytes):
    import ctypes
    from . import _lib as L
    from .engine import AssertFailed
    proof = ctypes.create_string_buffer(256)
    status = ctypes.c_int32(-1)
    err = ctypes.create_string_buffer(L.ERRCAP)
    rc = L.zke_shard_combine(zkey.handle, partials, world, rs, proof, ctypes.byref(status), err, L.ERRCAP)
    if rc < 0:
        raise L.ZkeError(err.value.decode())
    if rc > 0:
        raise AssertFailed(err.value.decode())
    return proof.raw, publics, status.value
