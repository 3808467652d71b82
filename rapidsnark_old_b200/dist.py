"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), MSMs sharded by point range.

Two exchange steps:
  * the H pipeline is three independent transform chains (a, b, c): with N > 1 each chain runs on ONE rank
    (`poly_owners`) and its result - domain_size x 32 bytes of coset evaluations - is broadcast to the others
    (NCCL broadcast over NVSwitch, enqueued behind the producing kernels on the library's H stream), instead of
    every rank repeating all three chains;
  * the per-GPU partial results of the five MSMs: 768 bytes per rank, one all_gather, then a handful of group
    additions on the host (NCCL has no elliptic-curve reduction operator).
Works with the gloo backend on CPU tensors too (tests)."""
import torch
import torch.distributed as dist

from . import fold_partials, groth16_finalize, groth16_finalize_prepared


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


def poly_owners(world):
    """Rank that runs the transform chain of a, b, c: spread over the first three ranks (all on rank 0 if alone)."""
    return [i % world for i in range(3)]


def poly_mask(rank, world):
    return sum(1 << i for i, o in enumerate(poly_owners(world)) if o == rank)


class _DeviceBytes:
    """n bytes of device memory owned by the library, viewed by torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


_views = {}   # torch views of the current zkey's a, b, c buffers + its H stream (rebuilt when the zkey changes)


def exchange_polys(tensors, group=None):
    """tensors = [a, b, c] (same shape on every rank; rank poly_owners()[i] holds the valid copy of tensors[i]):
    one broadcast per polynomial from its owner.  In place; afterwards every rank holds all three."""
    world = dist.get_world_size(group)
    works = [dist.broadcast(t, src=o, group=group, async_op=True) for t, o in zip(tensors, poly_owners(world))]
    for w in works:
        w.wait()


def prove_msms_distributed(zk, wtns, on_device, domain_size, device, group=None):
    """This rank's 768-byte partial record with the H pipeline spread over the ranks (see module docstring).
    `zk` is a rapidsnark_old_b200.ZKey uploaded with shard_index = rank, shard_count = world."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return zk.prove_msms_dev(wtns) if on_device else zk.prove_msms(wtns)
    rank = dist.get_rank(group)
    bufs, hstream = zk.prove_begin(wtns, on_device, poly_mask(rank, world))
    key = (tuple(bufs), hstream, domain_size)
    if key not in _views:
        _views.clear()
        _views[key] = ([torch.as_tensor(_DeviceBytes(p, domain_size * 32), device=device) for p in bufs],
                       torch.cuda.ExternalStream(hstream, device=device))
    views, ext = _views[key]
    with torch.cuda.stream(ext):
        exchange_polys(views, group)       # ordered after the transform kernels and before the combine, no host sync
    return zk.prove_finish()


def finish_proof(part768, vk, r32, s32, device=None, group=None, prep640=None):
    """partial MSM results of this rank -> (folded 768-byte record, proof A|B|C 256 bytes) on every rank.
    prep640: result of groth16_blind_prepare(vk, r32, s32) computed on a host thread while the GPU worked."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        part768 = fold_partials(all_gather_partials(part768, device, group))
    if prep640 is not None:
        return part768, groth16_finalize_prepared(part768, vk, prep640, r32, s32)
    return part768, groth16_finalize(part768, vk, r32, s32)
