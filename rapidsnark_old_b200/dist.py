"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), MSMs sharded by point range.

Two exchange steps:
  * the H pipeline is three independent transform chains (a, b, c): with N > 1 each chain runs on ONE rank
    (`poly_owners`); rank r then needs, of each polynomial, only the slice of coset evaluations that its part of h
    is combined from - the owner sends every other rank its slice (NCCL point-to-point over NVSwitch, enqueued
    behind the producing kernels on the library's H stream), instead of every rank repeating all three chains;
  * the per-GPU partial results of the five MSMs: 768 bytes per rank, one all_gather, then a handful of group
    additions on the host (NCCL has no elliptic-curve reduction operator).
Works with the gloo backend on CPU tensors too (tests)."""
import torch
import torch.distributed as dist

from . import fold_partials, groth16_finalize, groth16_finalize_prepared


def shard_range(length, index, count):
    """Index range [lo, hi) of a table of `length` points owned by shard `index` (same rule as b200_zkey_upload)."""
    return length * index // count, length * (index + 1) // count


# One transform chain of the H pipeline (its share of the a, b, c build, iNTT, twist, NTT) relative to the five MSMs of
# the whole proof.  A rank that runs a chain gets that much less of the MSMs, so that all ranks finish together.
# Alone a chain is 0.6 ms against 18.4 ms of MSMs at 2^20 on B200 (0.033: balances two ranks); next to the short MSMs
# of four or eight shards it costs about twice that, measured by timing an owner rank and a non-owner rank of the
# same world on one GPU (bench.py --emulate-shards N --emulate-rank R, profiles/r02_scaling_notes.md).
CHAIN_COST = {2: 0.033}
CHAIN_COST_DEFAULT = 0.045
PLAN_DEN = 1 << 16


def shard_plan(world, chain_cost=None):
    """[(lo_num, hi_num)] per rank in units of 1 / PLAN_DEN: the point range [len * lo / DEN, len * hi / DEN) of every
    table that rank owns.  Even split, minus chain_cost per transform chain the rank runs (poly_owners)."""
    if world == 1:
        return [(0, PLAN_DEN)]
    if chain_cost is None:
        chain_cost = CHAIN_COST.get(world, CHAIN_COST_DEFAULT)
    owners = poly_owners(world)
    share = [(1.0 + chain_cost * len(owners)) / world - chain_cost * owners.count(r) for r in range(world)]
    if min(share) <= 0:
        share = [1.0 / world] * world
    bounds, acc = [0], 0.0
    for r in range(world):
        acc += share[r]
        bounds.append(min(PLAN_DEN, int(round(acc * PLAN_DEN))))
    bounds[-1] = PLAN_DEN
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def plan_range(length, index, world, plan=None):
    """Index range of a table of `length` points owned by rank `index` under shard_plan (same rule as the library's
    explicit bounds: floor(length * num / den))."""
    lo, hi = (plan or shard_plan(world))[index]
    return length * lo // PLAN_DEN, length * hi // PLAN_DEN


_gather_bufs = {}   # (device, world) -> preallocated staging: pinned host in / out, device in / out
TIMES = {}          # B200_TIMING_PY=1: accumulated host seconds per step of the N > 1 path (bench.py reports them)


def _tick(name, t0):
    import time
    t1 = time.perf_counter()
    TIMES[name] = TIMES.get(name, 0.0) + (t1 - t0)
    return t1


def all_gather_partials(part768, device=None, group=None):
    """768-byte partial records of all ranks, in rank order.  On a GPU: one H2D from a pinned buffer, one
    all_gather_into_tensor (NCCL), one D2H into a pinned buffer - all preallocated."""
    world = dist.get_world_size(group)
    n = len(part768)
    if device is None or torch.device(device).type != "cuda":
        mine = torch.frombuffer(bytearray(part768), dtype=torch.uint8)
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine, group=group)
        return [bytes(t.numpy()) for t in outs]
    key = (str(device), world, n)
    if key not in _gather_bufs:
        _gather_bufs[key] = (torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n * world, dtype=torch.uint8).pin_memory(),
                             torch.empty(n, dtype=torch.uint8, device=device), torch.empty(n * world, dtype=torch.uint8, device=device))
    h_in, h_out, d_in, d_out = _gather_bufs[key]
    h_in.numpy()[:] = memoryview(part768)
    d_in.copy_(h_in, non_blocking=True)
    dist.all_gather_into_tensor(d_out, d_in, group=group)
    h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    raw = h_out.numpy().tobytes()
    return [raw[i * n:(i + 1) * n] for i in range(world)]


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


def exchange_polys(tensors, group=None, plan=None):
    """tensors = [a, b, c]: byte views of domain_size x 32 bytes, the same shape on every rank; rank poly_owners()[i]
    holds the valid copy of tensors[i].  Rank r combines only the slice plan_range(domain_size, r, world) of h (its
    range of the H table; plan = the shard_plan the zkeys were uploaded with, default shard_plan(world)), so of every polynomial it receives just that slice from the owner - point-to-point sends
    over NVLink, 1/world of a broadcast's volume per receiver.  In place."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = tensors[0].numel() // 32
    ops = []
    for t, o in zip(tensors, poly_owners(world)):
        peer = o if group is None else dist.get_global_rank(group, o)
        if rank == o:
            for k in range(world):
                lo, hi = plan_range(n, k, world, plan)
                if k != o and hi > lo:
                    dst = k if group is None else dist.get_global_rank(group, k)
                    ops.append(dist.P2POp(dist.isend, t[lo * 32:hi * 32], dst, group))
        else:
            lo, hi = plan_range(n, rank, world, plan)
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, t[lo * 32:hi * 32], peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def prove_msms_distributed(zk, wtns, on_device, domain_size, device, group=None, plan=None):
    """This rank's 768-byte partial record with the H pipeline spread over the ranks (see module docstring).
    `zk` is a rapidsnark_old_b200.ZKey uploaded with shard_index = rank, shard_count = world and
    shard_bounds = shard_plan(world)[rank] + (PLAN_DEN,) (or `plan`, the same list on every rank)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return zk.prove_msms_dev(wtns) if on_device else zk.prove_msms(wtns)
    import time
    rank = dist.get_rank(group)
    t = time.perf_counter()
    bufs, hstream = zk.prove_begin(wtns, on_device, poly_mask(rank, world))
    t = _tick("prove_begin (enqueue)", t)
    key = (tuple(bufs), hstream, domain_size)
    if key not in _views:
        _views.clear()
        _views[key] = ([torch.as_tensor(_DeviceBytes(p, domain_size * 32), device=device) for p in bufs],
                       torch.cuda.ExternalStream(hstream, device=device))
    views, ext = _views[key]
    with torch.cuda.stream(ext):
        exchange_polys(views, group, plan)       # ordered after the transform kernels and before the combine, no host sync
    t = _tick("exchange (enqueue)", t)
    part = zk.prove_finish()
    _tick("prove_finish (enqueue H MSM, wait, collect)", t)
    return part


def finish_proof(part768, vk, r32, s32, device=None, group=None, prep640=None):
    """partial MSM results of this rank -> (folded 768-byte record, proof A|B|C 256 bytes) on every rank.
    prep640: result of groth16_blind_prepare(vk, r32, s32) computed on a host thread while the GPU worked."""
    import time
    t = time.perf_counter()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = all_gather_partials(part768, device, group)
        t = _tick("all_gather of the partial records", t)
        part768 = fold_partials(parts)
    if prep640 is not None:
        out = part768, groth16_finalize_prepared(part768, vk, prep640, r32, s32)
    else:
        out = part768, groth16_finalize(part768, vk, r32, s32)
    _tick("fold + finalize (host)", t)
    return out
