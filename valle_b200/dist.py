"""Data-parallel sharding of independent utterances over the GPUs of one box.

Utterances are independent (the reference decodes them one at a time, valle/bin/infer.py:223-258),
so the path shards with NO data-path collective: rank r decodes a contiguous slice of the prompt
list on its own replica of the weights.  The only exchange is ONE all-gather of the final code
matrices (`[B_local, T_max, Q]` int16 + lengths) per batch -- NCCL over NVLink on GPUs, gloo in the
CPU tests.  SURVEY.md section 8e.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced slice [lo, hi) of n items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_codes(codes: Sequence[torch.Tensor], t_max: int, n_q: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """list of [T_b, Q] int64 -> ([B, t_max, Q] int16 zero-padded, [B] int32 lengths) on `device`."""
    B = len(codes)
    out = torch.zeros((B, t_max, n_q), dtype=torch.int16, device=device)
    lens = torch.zeros(B, dtype=torch.int32, device=device)
    for b, c in enumerate(codes):
        out[b, : c.shape[0]] = c.to(device=device, dtype=torch.int16)
        lens[b] = c.shape[0]
    return out, lens


def gather_codes(codes: Sequence[torch.Tensor], n_q: int, device, b_max: int = None,
                 t_max: int = None) -> List[torch.Tensor]:
    """All ranks receive the code matrices of every utterance, in global (rank-major) order.
    One all-gather of lengths (to size the padded block) + one of the payload."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [c for c in codes]
    world = dist.get_world_size()
    mine = torch.tensor([len(codes), max([c.shape[0] for c in codes], default=0)], dtype=torch.int32, device=device)
    sizes = torch.empty(2 * world, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(sizes, mine)
    sizes = sizes.view(world, 2).cpu()
    bm = int(sizes[:, 0].max()) if b_max is None else b_max
    tm = int(sizes[:, 1].max()) if t_max is None else t_max
    block, lens = pack_codes(codes, tm, n_q, device)
    if block.shape[0] < bm:
        block = torch.cat([block, torch.zeros((bm - block.shape[0], tm, n_q), dtype=block.dtype, device=device)])
        lens = torch.cat([lens, torch.zeros(bm - lens.shape[0], dtype=lens.dtype, device=device)])
    assert n_q % 2 == 0, "codes travel as int16 pairs viewed as int32 (gloo has no int16 collectives)"
    all_blocks = torch.empty((world * bm, tm, n_q), dtype=torch.int16, device=device)
    all_lens = torch.empty(world * bm, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(all_blocks.view(torch.int32), block.contiguous().view(torch.int32))
    dist.all_gather_into_tensor(all_lens, lens.contiguous())
    all_lens_h = all_lens.cpu()
    out = []
    for r in range(world):
        for b in range(int(sizes[r, 0])):
            i = r * bm + b
            out.append(all_blocks[i, : int(all_lens_h[i])].to(torch.int64))
    return out
