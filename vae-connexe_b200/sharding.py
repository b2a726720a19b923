"""Head sharding over ranks: rank r of R owns the contiguous head block
[lo, hi); the batch, the encoder and the noise are replicated.  Pure host logic + the two
collectives the semantics force (sum of dz over ranks lives in engine.backward; the GC row
all-gather and the scalar loss all-reduce are here).  Works on any torch.distributed backend, so
it is covered by world_size-2 gloo tests on CPU."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def head_range(p: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Balanced contiguous partition: the first p % R ranks get one extra head."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(p, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def _stage_through_host(t: torch.Tensor, group) -> bool:
    """gloo moves CUDA tensors for all_reduce / broadcast only: other collectives go through the host (this is the
    two-ranks-on-one-GPU test configuration, tests/gpu_dist_parity.py; NCCL never takes this path)."""
    return t.is_cuda and dist.get_backend(group) == "gloo"


def allgather_rows(local_rows: torch.Tensor, p: int, rank: int, world_size: int, group=None) -> torch.Tensor:
    """Stack every rank's [P_r, ...] row block into the full [p, ...] tensor (GC rows, per-head
    losses, predictions).  Shards may differ by one row, so blocks are padded to the widest."""
    widest = (p + world_size - 1) // world_size
    pad = torch.zeros((widest,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    if _stage_through_host(pad, group):
        host = pad.cpu()
        hb = [torch.empty_like(host) for _ in range(world_size)]
        dist.all_gather(hb, host, group=group)
        bufs = [b.to(pad.device) for b in hb]
    else:
        bufs = [torch.empty_like(pad) for _ in range(world_size)]
        dist.all_gather(bufs, pad, group=group)
    out = []
    for r in range(world_size):
        lo, hi = head_range(p, r, world_size)
        out.append(bufs[r][: hi - lo])
    return torch.cat(out, 0)


def allreduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class TorchComm:
    """The data-path collective of a head shard (sum of dz = sum_heads dh0 over ranks) through
    torch.distributed (NCCL over NVLink on the GPU box; capturable in a CUDA graph)."""

    def __init__(self, group):
        self.group = group

    def allreduce_dz(self, dz_part: torch.Tensor) -> None:
        dist.all_reduce(dz_part, op=dist.ReduceOp.SUM, group=self.group)

    def close(self) -> None:
        return None


class SymmComm:
    """dz all-reduce as ONE kernel over NVLink peer memory (csrc/dz_allreduce.cu): local head sum + one-shot all-reduce +
    latent backward.  torch.distributed._symmetric_memory supplies the plumbing only -- a buffer with the same layout on
    every rank and the peers' mappings of it; the exchange itself is P2P stores / loads issued by our kernel."""

    fused = True

    def __init__(self, k, group, B: int, Z: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.k, self.group = k, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        n = k.dz_allreduce_bytes(B, Z, self.world) // 4
        self.buf = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.ptrs = [int(x) for x in self.hdl.buffer_ptrs]
        self.buf.zero_()                                   # flags / epochs start at zero on every rank ...
        torch.cuda.synchronize(device)
        dist.barrier(group)                                # ... before any rank's first kernel can signal a peer
        self.B, self.Z = B, Z

    def latent_bwd(self, dh0, P, lat, eps, beta, kl_form, dlat, B):
        self.k.dz_allreduce_latent_bwd(dh0, P, self.ptrs, self.rank, self.world, lat, eps, beta, kl_form, dlat, None, B, self.Z)

    def allreduce_dz(self, dz_part: torch.Tensor) -> None:
        dist.all_reduce(dz_part, op=dist.ReduceOp.SUM, group=self.group)

    def close(self) -> None:
        self.hdl = None
        self.buf = None


def make_dz_comm(k, group, B: int, Z: int, device, current):
    """Upgrade the NCCL dz all-reduce to the fused peer-memory kernel when the ranks sit on NVLink-connected GPUs under the
    NCCL backend; every rank takes the same decision (a failed rendezvous on any rank keeps NCCL everywhere)."""
    import os
    if not isinstance(current, TorchComm) or os.environ.get("CRVAE_DZ", "symm") != "symm" or not hasattr(k, "dz_allreduce_latent_bwd"):
        return current
    if torch.device(device).type != "cuda" or dist.get_backend(group) != "nccl":
        return current
    comm, ok = None, 1.0
    try:
        comm = SymmComm(k, group, B, Z, device)
    except Exception as exc:                               # symmetric memory unavailable (no P2P / fabric): keep NCCL
        ok = 0.0
        import warnings
        warnings.warn(f"symmetric-memory dz all-reduce unavailable ({exc!r}); using NCCL")
    flag = torch.tensor([ok], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return comm if float(flag) == 1.0 else current
