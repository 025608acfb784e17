"""Data-parallel sharding of independent utterances over the GPUs of one box.

Utterances are independent (the reference decodes them one at a time, valle/bin/infer.py:223-258),
so the path shards with NO data-path collective: rank r decodes a contiguous slice of the prompt
list on its own replica of the weights.  The only exchange is ONE all-gather of the final code
matrices (`[B_local, T_max, Q]` int16 + lengths) per batch -- NCCL over NVLink on GPUs, gloo in the
CPU tests.  SURVEY.md section 8e.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced slice [lo, hi) of n items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_codes(codes: Sequence[torch.Tensor], n_q: int, device, b_max: Optional[int] = None,
                 g_max: Optional[int] = None, packed: Optional[torch.Tensor] = None, return_base: bool = False):
    """All ranks receive the code matrices of every utterance, in global (rank-major) order.

    ONE fixed-shape collective per batch: every rank contributes an int32 block
        [ n_utt | len_0 .. len_{b_max-1} | its packed [g_max, Q] int16 codes viewed as int32 ]
    (codes are < 1025, int16 pairs travel as int32 because gloo has no int16 collectives).  `codes` is the rank's
    list of [T_b, Q] matrices; `packed` optionally the same rows already concatenated on the device (the engine's
    own [G, Q] output tensor, so nothing is re-packed).  With `b_max` (utterances per rank) and `g_max` (frames per
    rank) given -- known up front for a fixed workload, e.g. bench.py -- nothing is exchanged or synchronised before
    the payload collective; otherwise one small all-gather of the two sizes precedes it.  Unpacking is one
    device-to-host copy of the length headers and views of one int64 tensor: no per-utterance kernels."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [c for c in codes]
    assert n_q % 2 == 0, "codes travel as int16 pairs viewed as int32 (gloo has no int16 collectives)"
    world = dist.get_world_size()
    n_mine = len(codes)
    lens_h = [int(c.shape[0]) for c in codes]
    g_mine = sum(lens_h)
    if b_max is None or g_max is None:
        mine = torch.tensor([n_mine, g_mine], dtype=torch.int32, device=device)
        sizes = torch.empty(2 * world, dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(sizes, mine)
        sizes = sizes.view(world, 2).cpu()
        b_max, g_max = int(sizes[:, 0].max()), int(sizes[:, 1].max())
    assert n_mine <= b_max and g_mine <= g_max, (n_mine, b_max, g_mine, g_max)
    hdr = 1 + b_max
    words = hdr + g_max * n_q // 2
    block = torch.zeros(words, dtype=torch.int32, device=device)
    block[:1 + n_mine] = torch.tensor([n_mine] + lens_h, dtype=torch.int32).to(device, non_blocking=True)
    if g_mine:
        if packed is None:
            packed = torch.cat([c.reshape(-1, n_q) for c in codes]) if n_mine > 1 else codes[0].reshape(-1, n_q)
        assert packed.shape == (g_mine, n_q)
        block[hdr: hdr + g_mine * n_q // 2] = packed.to(device=device, dtype=torch.int16).contiguous().view(torch.int32).view(-1)
    everything = torch.empty((world, words), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(everything.view(-1), block)
    heads = everything[:, :hdr].cpu()                                      # the one host sync
    allc = everything[:, hdr:].contiguous().view(torch.int16).view(world, g_max, n_q).to(torch.int64)
    out = []
    for r in range(world):
        off = 0
        for b in range(int(heads[r, 0])):
            n = int(heads[r, 1 + b])
            out.append(allc[r, off: off + n])
            off += n
    # return_base: also the backing [world, g_max, Q] int64 tensor the views point into (one D2H moves everything)
    return (out, allc) if return_base else out


def tokenize_sharded(extractor, samples: Sequence, sampling_rate: int) -> List[torch.Tensor]:
    """Dataset-scale EnCodec tokenisation over the GPUs of one box (the reference's single-GPU
    `compute_and_store_features_batch` path, valle/bin/tokenizer.py:172-214, README.md:144 TODO): rank r encodes a
    contiguous shard of the waveform list with its own replica of the codec (no data-path collective), then ONE
    all-gather returns every utterance's [T, 8] codes to every rank, in the original order."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_range(len(samples), rank, world)
    mine = extractor.extract_batch_device(samples[lo:hi], sampling_rate)
    if world == 1:
        return mine
    return gather_codes(mine, extractor.config.num_quantizers, extractor.tokenizer.device)
