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
