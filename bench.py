#!/usr/bin/env python
"""bench.py - Groth16 proof time (ms) at 2^20 constraints on B200 (BASELINE.json metric, configs[1]).

  python bench.py --gpus N --steps K --warmup W            own arm (one rank per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU prover on the host cores

One step = one whole proof of the synthetic 2^20 BN254 circuit: H pipeline
(a,b,c build, 3 x iNTT/twist/NTT, h) + five MSMs (H, A, B1 over G1; B2 over G2; C over G1) + blinding +
affine conversion.  With N > 1 every table is sharded by point range over the ranks (strong scaling: the
same proof, less work per GPU), each rank recomputes the (cheap) H pipeline, the 768-byte partial results
are exchanged with one NCCL all_gather and folded on the host.

  value : ms per proof, witness already resident in HBM            (higher_is_better = false)
  e2e   : ms per proof through the C-ABI with the witness in pinned HOST memory (H2D inside the timed region,
          proof bytes read back) - the headline number to compare with --impl reference
Timed with CUDA events recorded on the library's own stream (b200_stream), barrier + synchronize on both
sides, max over ranks.  Inputs (point tables 0.4 GB + coefficients 0.14 GB per proof) exceed the 126 MB L2.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before any CUDA call (see rapidsnark_old_b200/__init__.py)

METRIC = "groth16_proof_ms_2^20_constraints"   # --log-n K renames it to ..._2^K_...
UNIT = "ms"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- workload
def build_inputs(log_n, seed, g1_many, g2_many, fast=True, shard=None, bounds=None):
    """The synthetic circuit of SURVEY.md Appendix C.  fast: scalars from the library's host routine (FastSynth,
    identical values, seconds instead of minutes at 2^24); the point makers must then accept packed bytes.
    shard = (rank, world): only this rank's slices of the point tables are generated (s.shard_points)."""
    from rapidsnark_old_b200 import synth
    t = time.time()
    s = synth.FastSynth(log_n, seed) if fast else synth.Synth(log_n, seed)
    log("[bench] synthetic circuit 2^%d: scalars in %.1fs" % (log_n, time.time() - t))
    t = time.time()
    if shard:
        s.build_shard_points(g1_many, g2_many, *shard, bounds=bounds)
    else:
        s.build_points(g1_many, g2_many)
    log("[bench] point tables in %.1fs" % (time.time() - t))
    return s


def gpu_point_makers(ctx):
    from rapidsnark_old_b200 import synth
    g1, g2 = synth.g1_gen_bytes(), synth.g2_gen_bytes()
    return (lambda ks: ctx.fixed_base_g1(g1, synth.le32_many(ks), synth.count32(ks)),
            lambda ks: ctx.fixed_base_g2(g2, synth.le32_many(ks), synth.count32(ks)))


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def blinding_factors():
    """r, s of groth16.cpp:213-217: 31 random bytes each, top byte zero - fixed here so that runs are comparable, but
    FULL SIZE (248 bits): the blinding scalar multiplications are real work of every proof."""
    import hashlib
    return hashlib.sha256(b"b200 bench r").digest()[:31] + b"\0", hashlib.sha256(b"b200 bench s").digest()[:31] + b"\0"


def field_backend(o):
    """Which Montgomery product the CPU oracle runs: ADX assembly like the reference's generated routines, or C."""
    try:
        return "ADX assembly (mulx/adcx/adox)" if o.lib.oracle_field_backend() else "plain C (__int128 CIOS)"
    except Exception:
        return "plain C (__int128 CIOS)"


# ----------------------------------------------------------------------------- reference arm
class DiskInputs:
    """The circuit written by tools/make_inputs.py (a separate process), read back without loading libb200snark.so."""

    def __init__(self, d):
        import numpy as np
        self.meta = m = json.load(open(os.path.join(d, "meta.json")))
        self.n, self.n_vars, self.n_public, self.n_coefs, self.log_n = m["n"], m["n_vars"], m["n_public"], m["n_coefs"], m["log_n"]
        self._arr = {k: np.fromfile(os.path.join(d, k + ".bin"), dtype=np.uint8) for k in ("A", "B1", "B2", "C", "H", "coefs", "wtns")}
        self.vk = {k: bytes.fromhex(v) for k, v in m["vk"].items()}

    def ptr(self, k):
        return ctypes.c_void_p(self._arr[k].ctypes.data)


def reference_inputs(log_n):
    """Inputs for the reference arm, generated by a child process (GPU fixed-base when a device is there, CPU oracle
    otherwise) and cached under B200_BENCH_CACHE (default /tmp/b200_bench_inputs)."""
    d = os.path.join(os.environ.get("B200_BENCH_CACHE", "/tmp/b200_bench_inputs"), "chain_2e%d_seed2" % log_n)
    if not os.path.exists(os.path.join(d, "meta.json")):
        env = dict(os.environ)
        env.pop("OMP_NUM_THREADS", None)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_inputs.py"), "--log-n", str(log_n),
                               "--seed", "2", "--out", d], env=env)
    return DiskInputs(d)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm uses every host core it can, as the
    # reference's own build does (tasksfile.js:83 -fopenmp).  Set before libgomp is loaded AND through omp_set_num_threads.
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.ref()
    kind = "reference"
    if o is None:
        o, kind = oracle_lib.port(), "port"
    try:
        o.fn("set_threads")(ctypes.c_int(ncores))
    except AttributeError:
        pass
    log_n = args.log_n
    s = reference_inputs(log_n)
    vk = s.vk
    r32, s32 = blinding_factors()
    cores = o.threads()
    u32, u64 = ctypes.c_uint32, ctypes.c_uint64
    hoisted = kind == "reference" and hasattr(o.lib, "ref_make_prover")
    prover = None
    t_make = 0.0
    if hoisted:
        # Groth16::makeProver once (FFT root table of 2n entries, fft.cpp:32-115), Prover::prove per step: the same
        # split as this repo's b200_zkey_upload / prove, and what the reference's server mode does (fullprover.cpp:43-59)
        o.lib.ref_make_prover.restype = ctypes.c_void_p
        t0 = time.perf_counter()
        prover = ctypes.c_void_p(o.lib.ref_make_prover(u32(s.n_vars), u32(s.n_public), u32(s.n), u64(s.n_coefs),
                                                       vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"],
                                                       s.ptr("coefs"), s.ptr("A"), s.ptr("B1"), s.ptr("B2"), s.ptr("C"),
                                                       s.ptr("H")))
        t_make = time.perf_counter() - t0

    def step():
        out = ctypes.create_string_buffer(256)
        if hoisted:
            o.lib.ref_prover_prove(prover, s.ptr("wtns"), out)      # the untouched Prover::prove (r, s from randombytes)
            return out.raw
        m = ctypes.create_string_buffer(768)
        o.fn("prove_msms")(u32(s.n_vars), u32(s.n_public), u32(s.n), u64(s.n_coefs), s.ptr("coefs"), s.ptr("A"),
                           s.ptr("B1"), s.ptr("B2"), s.ptr("C"), s.ptr("H"), s.ptr("wtns"), m)
        return o.blind(m.raw, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], r32, s32)

    # bounded run: the reference takes seconds per proof; cap warm-up + steps so the arm ends within minutes
    t0 = time.perf_counter(); step(); first = time.perf_counter() - t0
    budget = 150.0
    steps = max(1, min(args.steps, int(budget / max(first, 1e-3)) - 1))
    warm = 0 if first > 5 else min(args.warmup, 1)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ms = (time.perf_counter() - t0) / steps * 1e3
    if hoisted:
        o.lib.ref_free_prover(prover)
    sample = ("full 2^%d proof (Prover::prove only, makeProver %.0f ms outside), %d timed run(s) after %d warm-up "
              "(first run %.0f ms)" % (log_n, t_make * 1e3, steps, warm + 1, first * 1e3)) if hoisted else \
             ("full 2^%d proof, %d timed run(s) after %d warm-up (first run %.0f ms)" % (log_n, steps, warm + 1, first * 1e3))
    loaded = [l.split()[-1] for l in open("/proc/self/maps") if "libb200snark" in l]
    line = {"impl": "reference", "metric": METRIC, "value": round(ms, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm + 1, "ms_per_step": round(ms, 3), "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "u256-mont", "data": "synthetic",
            "config": {"workload": "groth16 prove, BN254, 2^%d constraints, synthetic chain circuit" % log_n,
                       "n_vars": s.n_vars, "n_public": s.n_public, "n_coefs": s.n_coefs,
                       "impl_detail": ("reference templates (curve/multiexp/fft/groth16) compiled from /root/reference "
                                       "over a restated Fq/Fr field, OpenMP" if kind == "reference" else "plain-C oracle port") +
                                      "; field product: " + field_backend(o),
                       "inputs": "written by a child process (%s); libb200snark.so loaded in this process: %s"
                                 % (s.meta.get("tables_by"), "yes" if loaded else "no")},
            "cpu_baseline": {"value": round(ms, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(ms, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- own arm
def run_own(args):
    import torch
    import torch.distributed as dist
    import rapidsnark_old_b200 as b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b200.Context(local)
    log_n = args.log_n
    if log_n >= 25:
        # 2^26: every rank holds ~10 GB of scalar vectors, chain owners another ~9 GB of coefficient records
        try:
            import psutil
            need = (12 + (10 if rank < 3 else 0)) * (1 << 30) * (1 << (log_n - 26)) * world // max(1, world // 8 or 1)
            avail = psutil.virtual_memory().available
            if rank == 0:
                log("[bench] host memory: %.0f GB available, about %.0f GB needed by %d ranks" % (avail / 2**30, (12 * world + 30) * 2.0 ** (log_n - 26), world))
            if avail < (12 * world + 30) * (1 << 30) * 2.0 ** (log_n - 26):
                if rank == 0:
                    print(json.dumps({"metric": "groth16_proof_ms_2^%d_constraints" % log_n, "skipped": "not enough host memory for the synthetic inputs"}), flush=True)
                return 0
        except ImportError:
            pass
    # big circuits on several GPUs: every rank generates only the table slices it uploads (2^26: 24 GB of tables)
    shard_inputs = world > 1 and (args.shard_inputs or log_n >= 23)
    from rapidsnark_old_b200 import dist as bdist
    # N > 1: ranks that run a transform chain of the H pipeline own a smaller point range (dist.shard_plan)
    plan = bdist.shard_plan(world) if not args.even_shards else [(bdist.PLAN_DEN * r // world, bdist.PLAN_DEN * (r + 1) // world) for r in range(world)]
    bounds = plan[rank] + (bdist.PLAN_DEN,)
    s = build_inputs(log_n, 2, *gpu_point_makers(ctx), shard=(rank, world) if shard_inputs else None, bounds=bounds)
    if shard_inputs:
        p = {k: s.shard_table_address(k) for k in ("A", "B1", "B2", "C", "H")}
    else:
        p = s.points
    vk = s.vk
    emu = args.emulate_shards if world == 1 else 0      # tuning aid: this GPU plays rank 0 of `emu` (no collective,
    for kv in args.opt:                                  # result unchecked); never used for a reported line
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    if emu:   # rank --emulate-rank of `emu`, with that world's plan
        er = args.emulate_rank
        eplan = bdist.shard_plan(emu) if not args.even_shards else [(bdist.PLAN_DEN * r // emu, bdist.PLAN_DEN * (r + 1) // emu) for r in range(emu)]
        bounds = eplan[er] + (bdist.PLAN_DEN,)
        if args.emulate_poly_mask == -2:
            args.emulate_poly_mask = bdist.poly_mask(er, emu)
    # a rank that runs no transform chain never reads the coefficients: it uploads the zkey without section 4
    need_coefs = world == 1 or emu or args.replicate_h or bdist.poly_mask(rank, world) != 0
    coefs = (s.coefs_array() if hasattr(s, "coefs_array") else s.coefs_section()) if need_coefs else None   # 8.9 GB at 2^26: no copies
    zk = ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"],
                         p["H"], args.emulate_rank if emu else rank, emu if emu else world, shard_bounds=bounds if (world > 1 or emu) else None)
    wt_bytes = s.wtns_bytes()
    # witness: pinned host copy (e2e) and device copy (value)
    wt_host = torch.empty(len(wt_bytes), dtype=torch.uint8).pin_memory()
    wt_host.copy_(torch.frombuffer(bytearray(wt_bytes), dtype=torch.uint8))
    wt_dev = wt_host.cuda()
    r32, s32 = blinding_factors()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    from concurrent.futures import ThreadPoolExecutor
    host_pool = ThreadPoolExecutor(max_workers=1)

    def blind_prepare():
        # the part of the blinding that needs only the key and r, s (r*delta1, s*delta1, rs*delta1, s*delta2) runs on
        # a host thread while the GPU computes the MSMs - every proof, r and s are per-proof values
        return host_pool.submit(b200.groth16_blind_prepare, vk, r32, s32)

    def finish(part, prep):
        # N > 1: all_gather of the 768-byte partial records (NCCL), fold, then the rest of the blinding + to-affine
        return bdist.finish_proof(part, vk, r32, s32, device=torch.device("cuda", local), prep640=prep.result())

    dev = torch.device("cuda", local)
    spread_h = world > 1 and not args.replicate_h    # N > 1: a, b, c transform chains on different ranks + NCCL broadcasts

    fused = world == 1 and not emu      # one GPU: b200_groth16_prove, the whole proof in one C call

    def emu_spread(ptr, on_device):
        # tuning aid: rank 0 of `emu` shards with only the transform chains in --emulate-poly-mask run here and no
        # exchange (the other polynomials are whatever the buffers hold: the result is meaningless and unchecked)
        prep = blind_prepare()
        zk.prove_begin(ptr, on_device, args.emulate_poly_mask)
        return finish(zk.prove_finish(), prep)

    def step_resident():
        if emu and args.emulate_poly_mask >= 0:
            return emu_spread(wt_dev.data_ptr(), True)
        if fused:
            return zk.prove(wt_dev.data_ptr(), vk, r32, s32, on_device=True)
        prep = blind_prepare()
        if spread_h:
            return finish(bdist.prove_msms_distributed(zk, wt_dev.data_ptr(), True, s.n, dev, plan=plan), prep)
        return finish(zk.prove_msms_dev(wt_dev.data_ptr()), prep)

    def step_e2e():
        if emu and args.emulate_poly_mask >= 0:
            return emu_spread(wt_host.data_ptr(), False)
        if fused:
            return zk.prove(wt_host.data_ptr(), vk, r32, s32)
        prep = blind_prepare()
        if spread_h:
            return finish(bdist.prove_msms_distributed(zk, wt_host.data_ptr(), False, s.n, dev, plan=plan), prep)
        return finish(zk.prove_msms(wt_host.data_ptr()), prep)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness gate before timing: every result is checked against the known discrete logs
    msms, proof = step_e2e()
    if not emu:
        check_known_dlogs(b200, s, msms, proof, r32, s32)

    def timed(fn, steps, warm):
        """ms per step over `steps` timed steps (CUDA events on the library's stream, barrier + synchronize on both
        sides, max over ranks) - nothing but the product call inside the timed region - then the per-phase device
        times of three more, untimed, steps (reading the phase events back is instrumentation)."""
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        launches = (ctx.launch_count() - l0) // steps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        phases = {}
        for _ in range(3):
            fn()
            for k, v in ctx.phase_ms().items():
                phases[k] = phases.get(k, 0.0) + v / 3
        barrier()
        return ms, phases, launches

    sampler = ClockSampler(local)
    sampler.start()
    bdist.TIMES.clear()
    ms_res, ph_res, launches = timed(step_resident, args.steps, args.warmup)
    host_steps = {k: round(v * 1e3 / (args.steps + args.warmup + 3), 4) for k, v in bdist.TIMES.items()}   # N > 1: host ms per proof
    ms_e2e, ph_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop()

    # -- extra, shorter measurements of the same step (reported inside the same JSON line) -------------------------
    extra_steps = max(3, args.steps // 2)
    # (1) the witness in ordinary PAGEABLE host memory, as the CLI / FullProver hand it over (an mmap of the .wtns file)
    wt_pageable = ctypes.create_string_buffer(wt_bytes, len(wt_bytes))
    wt_pageable_addr = ctypes.addressof(wt_pageable)

    def step_e2e_pageable():
        if fused:
            return zk.prove(wt_pageable_addr, vk, r32, s32)
        prep = blind_prepare()
        if spread_h:
            return finish(bdist.prove_msms_distributed(zk, wt_pageable_addr, False, s.n, dev, plan=plan), prep)
        return finish(zk.prove_msms(wt_pageable_addr), prep)

    ms_pageable, _, _ = timed(step_e2e_pageable, extra_steps, 3)
    # (2) a circom-like witness (70 % of the wires in {0,1}, 20 % below 2^32, 10 % uniform)
    # on the same zkey; the uniform full-width witness above is the worst case for the MSMs.  It is not a satisfying
    # assignment - the five points are checked against the known discrete logs of the tables instead.
    circom = None
    if not emu and log_n <= 24:      # (the host-side check of this witness is Python big-integer work: minutes at 2^26)
        cw = circom_like_witness(s.n_vars)
        cw_host = torch.empty(len(cw), dtype=torch.uint8).pin_memory()
        cw_host.copy_(torch.frombuffer(bytearray(cw), dtype=torch.uint8))
        cw_dev = cw_host.cuda()

        def step_circom_resident():
            if fused:
                return zk.prove(cw_dev.data_ptr(), vk, r32, s32, on_device=True)
            prep = blind_prepare()
            if spread_h:
                return finish(bdist.prove_msms_distributed(zk, cw_dev.data_ptr(), True, s.n, dev, plan=plan), prep)
            return finish(zk.prove_msms_dev(cw_dev.data_ptr()), prep)

        def step_circom_e2e():
            if fused:
                return zk.prove(cw_host.data_ptr(), vk, r32, s32)
            prep = blind_prepare()
            if spread_h:
                return finish(bdist.prove_msms_distributed(zk, cw_host.data_ptr(), False, s.n, dev, plan=plan), prep)
            return finish(zk.prove_msms(cw_host.data_ptr()), prep)

        cm, _ = step_circom_e2e()
        check_witness_msms(b200, s, cw, cm)
        ms_c_res, _, _ = timed(step_circom_resident, extra_steps, 3)
        ms_c_e2e, _, _ = timed(step_circom_e2e, extra_steps, 3)
        circom = {"value": round(ms_c_res, 4), "e2e": round(ms_c_e2e, 4), "unit": UNIT, "steps": extra_steps,
                  "witness": "70% of wires in {0,1}, 20% < 2^32, 10% uniform (SURVEY 8d config 2); pi_a, pib1, pi_b, pi_c "
                             "checked against the tables' known discrete logs"}
    # (3) one more proof with the segment timeline on: which phase runs when, per stream (critical path)
    ctx.set_option("timeline", 1)
    step_resident()
    tl = ctx.timeline()
    ctx.set_option("timeline", 0)

    pk, pk_kind = peaks()
    # dominant kernel: k_msm_accumulate<Fq> - 4 launches per proof (H, A, B1, C); algorithmic bytes 96 B/point
    zl = lambda total: total * bounds[1] // bounds[2] - total * bounds[0] // bounds[2]     # this rank's share
    n_pts = [zl(s.n), zl(s.n_vars), zl(s.n_vars), zl(s.n_vars - s.n_public - 1)]
    # launches of the G1 accumulation kernel per proof: one fused launch for all four MSMs on a single GPU, two
    # (witness MSMs, then H) on sharded zkeys - counted from the timeline's segments
    tls = timeline_summary(tl)
    acc_launches = max(1, int(tls.get("msm_accumulate_g1", {}).get("segments", 4)))
    alg_bytes = 96.0 * sum(n_pts) / acc_launches
    acc_ms = ph_res.get("msm_accumulate_g1", 0.0) / acc_launches
    achieved = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic(log_n, world)
    roof = {"bound": "hbm", "kernel": "k_msm_accumulate_sets<Fq> (G1 bucket accumulation)", "achieved": round(achieved, 2), "peak": pk["hbm_gbs"],
            "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 5), "traffic": traffic, "traffic_source": traffic_src,
            "peak_kind": pk_kind, "launches_per_proof": acc_launches,
            "launch_ms": round(acc_ms, 4), "algorithmic_bytes_per_launch": alg_bytes,
            "note": "integer-ALU bound by construction (SURVEY 8d): ~10 Montgomery products of 8x8 32-bit limbs per "
                    "96 algorithmic bytes; see DESIGN.md for the IMAD roofline"}

    line = {"metric": METRIC, "value": round(ms_res, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_res, 4), "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "u256-mont", "data": "synthetic",
            "config": {"workload": "groth16 prove, BN254, 2^%d constraints, synthetic chain circuit" % log_n,
                       "n_vars": s.n_vars, "n_public": s.n_public, "n_coefs": s.n_coefs,
                       "parallelism": ("point-range shards x%d (shares %s), 1 NCCL all_gather of 768 B partials" %
                                       (world, [round((b - a) / bdist.PLAN_DEN, 4) for a, b in plan])) +
                                      (", H transform chains a/b/c on ranks %s, every rank receives its slice of each "
                                       "(NCCL send/recv, %d KB per slice)" %
                                       (bdist.poly_owners(world), (s.n * 32 // world) >> 10) if spread_h else ""),
                       "l2": "inputs larger than L2 (0.5 GB of tables and coefficients per proof)"},
            "e2e": {"value": round(ms_e2e, 4), "unit": UNIT, "h2d_bytes_per_step": len(wt_bytes), "d2h_bytes_per_step": 768,
                    "pageable_host_witness_ms": round(ms_pageable, 4)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "circom_like_witness": circom,
            # per-phase sums of stream time: phases on different streams OVERLAP (they add up to more than the step);
            # the timeline below is the schedule of one proof
            "phase_stream_ms": {k: round(v, 4) for k, v in ph_res.items()},
            "phase_stream_ms_e2e": {k: round(v, 4) for k, v in ph_e2e.items()},
            "timeline_ms": tls}
    if host_steps:
        line["host_path_ms"] = host_steps
    if emu:
        line["config"]["emulated_shards"] = emu
        line["metric"] += "_EMULATED_RANK0_OF_%d" % emu
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not emu:
        line["cpu_baseline"] = cpu_baseline(s)
    if rank == 0:
        print(json.dumps(line), flush=True)
    zk.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_traffic(log_n, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of k_msm_accumulate<Fq>, from the newest committed
    `ncu --set full` capture summary (profiles/r*_accumulate_traffic.json, written by tools/summarize_profiles.py
    together with the commit and configuration it was taken at).  A capture of another configuration is not this
    run's traffic: then the value is None and the source says why."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_accumulate_traffic.json")))
    if not files:
        return None, "no capture summary under profiles/"
    try:
        d = json.load(open(files[-1]))
        src = "%s (commit %s, 2^%s, %s GPU)" % (os.path.basename(files[-1]), d.get("commit"), d.get("log_n"), d.get("n_gpus"))
        if d.get("log_n") != log_n or d.get("n_gpus") != world:
            return None, "capture is of another configuration: " + src
        return int(d["dram_bytes_per_launch"]), src
    except Exception as e:
        return None, "unreadable capture summary: %s" % e


def timeline_summary(tl):
    """[(phase, start, end)] segments of one proof -> per phase [first start, last end, busy ms, segments], plus the
    step's span: a compact critical-path view (phases on different streams overlap)."""
    out = {}
    for name, t0, t1 in tl:
        o = out.setdefault(name, [t0, t1, 0.0, 0])
        o[0], o[1], o[2], o[3] = min(o[0], t0), max(o[1], t1), o[2] + (t1 - t0), o[3] + 1
    res = {k: {"start": round(v[0], 3), "end": round(v[1], 3), "busy": round(v[2], 3), "segments": v[3]} for k, v in out.items()}
    if tl:
        res["_span"] = round(max(t1 for _, _, t1 in tl), 3)
    return res


def circom_like_witness(n_vars, seed=7):
    import numpy as np
    rng = np.random.default_rng(seed)
    w = np.zeros((n_vars, 4), dtype=np.uint64)
    u = rng.random(n_vars)
    small = u < 0.7
    w[small, 0] = rng.integers(0, 2, size=int(small.sum()), dtype=np.uint64)
    mid = (u >= 0.7) & (u < 0.9)
    w[mid, 0] = rng.integers(0, 1 << 32, size=int(mid.sum()), dtype=np.uint64)
    wide = u >= 0.9
    w[wide] = rng.integers(0, 1 << 61, size=(int(wide.sum()), 4), dtype=np.uint64)   # < 2^253 < r
    w[0] = (1, 0, 0, 0)
    return w.tobytes()


def check_witness_msms(b200, s, wt, msms):
    """Any witness: pi_a, pib1, pi_b, pi_c are sum w_i * (table dlog_i) times the generator (the tables' discrete logs
    are known).  Computed with numpy object arithmetic - a few seconds at 2^20."""
    from rapidsnark_old_b200 import synth
    import numpy as np
    R = synth.R

    def ints(buf):
        a = np.frombuffer(buf, dtype="<u8").reshape(-1, 4).astype(object)
        return a[:, 0] + (a[:, 1] << 64) + (a[:, 2] << 128) + (a[:, 3] << 192)

    w = ints(wt)
    P = s.n_public
    ea = int((w * ints(s.A_tau)).sum() % R)
    eb = int((w * ints(s.B_tau)).sum() % R)
    ec = int((w[P + 1:] * ints(s.c_scalars)).sum() % R)
    g1, g2 = synth.g1_gen_bytes(), synth.g2_gen_bytes()
    le = lambda v: int(v).to_bytes(32, "little")
    assert b200.host_g1_to_affine(msms[128:256]) == b200.host_g1_to_affine(b200.host_g1_mul(g1, le(ea))), "pi_a"
    assert b200.host_g1_to_affine(msms[256:384]) == b200.host_g1_to_affine(b200.host_g1_mul(g1, le(eb))), "pib1"
    assert b200.host_g2_to_affine(msms[384:640]) == b200.host_g2_to_affine(b200.host_g2_mul(g2, le(eb))), "pi_b"
    assert b200.host_g1_to_affine(msms[640:768]) == b200.host_g1_to_affine(b200.host_g1_mul(g1, le(ec))), "pi_c"
    log("[bench] circom-like witness: pi_a, pib1, pi_b, pi_c match their known discrete logs")


def zk_len(total, rank, world):
    return total * (rank + 1) // world - total * rank // world


def check_known_dlogs(b200, s, msms, proof, r32, s32):
    """Size-independent correctness gate: with the toxic waste known, A, B, C of the proof are single scalar
    multiples of the generators and must satisfy the Groth16 equation."""
    from rapidsnark_old_b200 import synth
    R = synth.R
    ea, eb = s.dlog_a, s.dlog_b
    r, t = int.from_bytes(r32, "little"), int.from_bytes(s32, "little")
    a = (s.alpha + ea + r * s.delta) % R
    b = (s.beta + eb + t * s.delta) % R
    g1, g2 = synth.g1_gen_bytes(), synth.g2_gen_bytes()
    A = b200.host_g1_to_affine(b200.host_g1_mul(g1, a.to_bytes(32, "little")))
    B = b200.host_g2_to_affine(b200.host_g2_mul(g2, b.to_bytes(32, "little")))
    assert proof[:64] == A, "proof.A does not match its known discrete log"
    assert proof[64:192] == B, "proof.B does not match its known discrete log"
    # C: solve the verification equation for c and compare: a*b = alpha*beta + pub + c*delta
    c = (a * b - s.alpha * s.beta - s.dlog_pub) * pow(s.delta, -1, R) % R
    C = b200.host_g1_to_affine(b200.host_g1_mul(g1, c.to_bytes(32, "little")))
    assert proof[192:256] == C, "proof.C does not satisfy the Groth16 verification equation"
    log("[bench] proof verified in the exponent (A, B, C match; e(A,B) = e(alpha,beta) e(pub,gamma) e(C,delta))")


def cpu_baseline(s):
    """The reference's CPU prover (oracle/_ref when present, else the plain-C port) on this box's host cores:
    Groth16::makeProver once, then ONE full Prover::prove of the same inputs (a bounded sample: a 2^20 proof is
    seconds of CPU work on all cores)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.ref()
    kind = "reference"
    if o is None:
        o, kind = oracle_lib.port(), "port"
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        o.fn("set_threads")(ctypes.c_int(ncores))
    except AttributeError:
        pass
    p, vk = s.points, s.vk
    coefs, wt = s.coefs_section(), s.wtns_bytes()
    u32, u64 = ctypes.c_uint32, ctypes.c_uint64
    if kind == "reference" and hasattr(o.lib, "ref_make_prover"):
        o.lib.ref_make_prover.restype = ctypes.c_void_p
        pr = ctypes.c_void_p(o.lib.ref_make_prover(u32(s.n_vars), u32(s.n_public), u32(s.n), u64(s.n_coefs), vk["alpha1"],
                                                   vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], coefs, p["A"],
                                                   p["B1"], p["B2"], p["C"], p["H"]))
        out = ctypes.create_string_buffer(256)
        t0 = time.perf_counter()
        o.lib.ref_prover_prove(pr, wt, out)
        ms = (time.perf_counter() - t0) * 1e3
        o.lib.ref_free_prover(pr)
        what = "one full 2^%d Prover::prove (makeProver outside), same inputs, cold" % s.log_n
    else:
        t0 = time.perf_counter()
        m = o.prove_msms(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"], wt)
        o.blind(m, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], *blinding_factors())
        ms = (time.perf_counter() - t0) * 1e3
        what = "one full 2^%d proof (H pipeline + 5 MSMs + blinding), same inputs, cold" % s.log_n
    return {"value": round(ms, 1), "unit": UNIT, "cores": o.threads(), "kind": kind,
            "sample": "%s; field product: %s" % (what, field_backend(o))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-inputs", action="store_true",
                    help="N > 1: generate only this rank's slices of the point tables (automatic from 2^23)")
    ap.add_argument("--even-shards", action="store_true", help="N > 1: equal point ranges on every rank (A/B)")
    ap.add_argument("--replicate-h", action="store_true", help="N > 1: every rank runs the whole H pipeline (A/B)")
    ap.add_argument("--emulate-shards", type=int, default=0, help="tuning only: time rank 0 of K shards on one GPU")
    ap.add_argument("--emulate-rank", type=int, default=0, help="tuning only, with --emulate-shards: which rank to play")
    ap.add_argument("--emulate-poly-mask", type=int, default=-1,
                    help="tuning only, with --emulate-shards: run only these transform chains (bit 0 a, 1 b, 2 c; -2 = the "
                         "emulated rank's own), no exchange")
    ap.add_argument("--opt", nargs="*", default=[], help="tuning only: library options name=value (b200_set_option)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    global METRIC
    METRIC = "groth16_proof_ms_2^%d_constraints" % args.log_n
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
