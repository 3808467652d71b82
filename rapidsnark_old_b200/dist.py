"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), MSMs sharded by point range.

The only exchange step of the path is the combination of the per-GPU partial results of the five MSMs
(SURVEY.md 8e): 768 bytes per rank, one all_gather, then a handful of group additions on the host (NCCL has
no elliptic-curve reduction operator).  Works with the gloo backend on CPU tensors too (tests)."""
import torch
import torch.distributed as dist

from . import fold_partials, groth16_finalize


def shard_range(length, index, count):
    """Index range [lo, hi) of a table of `length` points owned by shard `index` (same rule as b200_zkey_upload)."""
    return length * index // count, length * (index + 1) // count


def all_gather_partials(part768, device=None, group=None):
    world = dist.get_world_size(group)
    mine = torch.frombuffer(bytearray(part768), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    return [bytes(t.cpu().numpy()) for t in outs]


def finish_proof(part768, vk, r32, s32, device=None, group=None):
    """partial MSM results of this rank -> (folded 768-byte record, proof A|B|C 256 bytes) on every rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        part768 = fold_partials(all_gather_partials(part768, device, group))
    return part768, groth16_finalize(part768, vk, r32, s32)
