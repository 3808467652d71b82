"""bench.py's own-arm control flow on the CPU: torch.cuda and the library context are replaced by stand-ins backed by
the oracle, so that the exact code the driver runs on the GPU box (input builder, upload call, step functions with the
host-thread blinding, the in-the-exponent proof check, timing loop, JSON line) is exercised in the CPU suite.
Nothing here measures anything."""
import ctypes
import io
import json
import sys
import types
from contextlib import redirect_stdout

import pytest

import oracle_lib
import synth_util

sys.path.insert(0, oracle_lib.ROOT)
import bench
import rapidsnark_old_b200 as b200


class _FakeZKey:
    def __init__(self, ctx, args):
        (self.n_vars, self.n_public, self.n, self.n_coefs, self.coefs, self.A, self.B1, self.B2, self.C, self.H,
         self.index, self.count) = args
        self.ctx = ctx
        if hasattr(self.coefs, "tobytes"):
            self.coefs = self.coefs.tobytes()       # bench hands the section over as a numpy byte buffer
        assert (self.index, self.count) == (0, 1)

    def _prove(self, ptr):
        wt = ctypes.string_at(ptr, self.n_vars * 32)
        self.ctx.launches += 60
        return self.ctx.o.prove_msms(self.n_vars, self.n_public, self.n, self.n_coefs, self.coefs, self.A, self.B1,
                                     self.B2, self.C, self.H, wt)

    prove_msms = _prove
    prove_msms_dev = _prove

    def prove(self, ptr, vk, r32, s32, on_device=False):
        m = self._prove(ptr)
        return m, self.ctx.o.blind(m, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], r32, s32)

    def free(self):
        pass


class _FakeContext:
    def __init__(self, device):
        self.o = oracle_lib.best()
        self.launches = 0
        g1m, g2m = synth_util.oracle_point_makers(self.o)
        unpack = lambda ks: [int.from_bytes(ks[i:i + 32], "little") for i in range(0, len(ks), 32)]
        self.fixed_base_g1 = lambda base, ks, n: g1m(unpack(ks))
        self.fixed_base_g2 = lambda base, ks, n: g2m(unpack(ks))

    def zkey_upload(self, *args, shard_bounds=None):
        assert shard_bounds is None            # one GPU: no explicit bounds
        return _FakeZKey(self, args)

    def set_option(self, k, v):
        pass

    def stream(self):
        return 0

    def launch_count(self):
        return self.launches

    def phase_ms(self):
        return {"msm_accumulate_g1": 4.0, "msm_accumulate_g2": 2.0}

    def timeline(self):
        return [("msm_sort", 0.0, 0.5), ("msm_accumulate_g2", 0.5, 2.5), ("ntt_h", 0.6, 2.0), ("msm_accumulate_g1", 2.5, 6.5)]

    def close(self):
        pass


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 10.0


def test_own_arm_flow_with_stand_ins(monkeypatch):
    import torch
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda *a, **k: object())
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(b200, "Context", _FakeContext)
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=3, impl="own", log_n=6, no_cpu_baseline=False, emulate_shards=0,
                                 opt=[], replicate_h=False, shard_inputs=False, emulate_poly_mask=-1, even_shards=False, emulate_rank=0)
    out = io.StringIO()
    with redirect_stdout(out):
        assert bench.run_own(args) == 0
    line = json.loads(out.getvalue().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["value"] == 5.0 and line["gpu_launches"] == 60 and line["higher_is_better"] is False
    assert line["e2e"]["h2d_bytes_per_step"] == (64 - 6) * 32 and line["e2e"]["d2h_bytes_per_step"] == 768
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_source"}
    assert line["roofline"]["traffic"] is None        # the committed capture is of 2^20, this run is 2^6
    assert line["circom_like_witness"]["value"] == 3.3333 and line["e2e"]["pageable_host_witness_ms"] == 3.3333   # 10 ms / 3 steps
    assert line["timeline_ms"]["_span"] == 6.5 and line["timeline_ms"]["ntt_h"]["busy"] == 1.4
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert "workload" in line["config"]
