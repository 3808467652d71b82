"""Golden vectors produced by the reference itself (tests/golden/make_golden.py -> golden_v1.json):
CPU checks of the oracle port and of the product's host side here; the GPU checks live at the bottom
(marked gpu) and include the CLI producing byte-identical proof.json / public.json."""
import ctypes
import json
import os
import subprocess

import pytest

import bn254 as bn
import oracle_lib
import rapidsnark_old_b200 as b200

ROOT = oracle_lib.ROOT
G = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_v1.json")))
C = G["circuit_2_4"]
PROVER = os.path.join(ROOT, "build", "prover")


def _sections(raw):
    """iden3 binfile -> {type: payload}"""
    n = int.from_bytes(raw[8:12], "little")
    pos, out = 12, {}
    for _ in range(n):
        typ = int.from_bytes(raw[pos:pos + 4], "little")
        size = int.from_bytes(raw[pos + 4:pos + 12], "little")
        out.setdefault(typ, raw[pos + 12:pos + 12 + size])
        pos += 12 + size
    return out


def _circuit():
    z = _sections(bytes.fromhex(C["zkey"]))
    w = _sections(bytes.fromhex(C["wtns"]))
    hdr = z[2]
    n_vars, n_public, domain = (int.from_bytes(hdr[72 + 4 * i:76 + 4 * i], "little") for i in range(3))
    vk = {"alpha1": hdr[84:148], "beta1": hdr[148:212], "beta2": hdr[212:340], "gamma2": hdr[340:468],
          "delta1": hdr[468:532], "delta2": hdr[532:660]}
    return dict(n_vars=n_vars, n_public=n_public, domain=domain, n_coefs=len(z[4]) // 44, coefs=z[4], A=z[5], B1=z[6],
                B2=z[7], C=z[8], H=z[9], vk=vk, wtns=w[2])


def _affine_to_xyzz(aff, g2=False):
    one = bn.to_mont(1)
    if g2:
        return aff if aff == bytes(128) and False else aff + (one + bytes(32)) * 2
    return aff + one * 2


# ----------------------------------------------------------------------------- CPU
def test_port_oracle_matches_reference_golden():
    o = oracle_lib.port()
    m = G["msm_g1"]
    assert o.g1_to_affine(o.g1_msm(bytes.fromhex(m["bases"]), bytes.fromhex(m["scalars"]), m["n"])).hex() == m["affine"]
    m = G["msm_g2"]
    assert o.g2_to_affine(o.g2_msm(bytes.fromhex(m["bases"]), bytes.fromhex(m["scalars"]), m["n"])).hex() == m["affine"]
    t = G["ntt"]
    assert o.fr_fft(bytes.fromhex(t["in"])).hex() == t["fft"]
    assert o.fr_ifft(bytes.fromhex(t["in"])).hex() == t["ifft"]
    c = _circuit()
    assert o.h_scalars(c["domain"], c["n_coefs"], c["coefs"], c["wtns"]).hex() == C["h_scalars"]
    msms = o.prove_msms(c["n_vars"], c["n_public"], c["domain"], c["n_coefs"], c["coefs"], c["A"], c["B1"], c["B2"],
                        c["C"], c["H"], c["wtns"])
    assert [a.hex() for a in o.msms_to_affine(msms)] == C["msms_affine"]
    vk = c["vk"]
    proof = o.blind(msms, vk["alpha1"], vk["beta1"], vk["beta2"], vk["delta1"], vk["delta2"], bytes.fromhex(C["r"]),
                    bytes.fromhex(C["s"]))
    assert proof.hex() == C["proof"]


def test_host_finalize_and_json_match_reference_cli_golden():
    c = _circuit()
    a = [bytes.fromhex(x) for x in C["msms_affine"]]
    msms = (_affine_to_xyzz(a[0]) + _affine_to_xyzz(a[1]) + _affine_to_xyzz(a[2]) + _affine_to_xyzz(a[3], True) +
            _affine_to_xyzz(a[4]))
    proof = b200.groth16_finalize(msms, c["vk"], bytes.fromhex(C["r"]), bytes.fromhex(C["s"]))
    assert proof.hex() == C["proof"]
    assert b200.proof_json(proof) == C["proof_json"]       # byte-identical to the reference CLI's proof.json


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    from rapidsnark_old_b200 import dist as bdist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    o = oracle_lib.port()
    c = _circuit()
    # this rank's point-range shard, computed by the oracle (no GPU here): same partition as b200_zkey_upload
    lo, hi = bdist.shard_range(c["n_vars"], rank, world)
    hlo, hhi = bdist.shard_range(c["domain"], rank, world)
    h = o.h_scalars(c["domain"], c["n_coefs"], c["coefs"], c["wtns"])
    w = c["wtns"]
    skip = c["n_public"] + 1
    clo, chi = max(lo, skip), max(hi, skip)
    part = (o.g1_msm(c["H"][64 * hlo:64 * hhi], h[32 * hlo:32 * hhi], hhi - hlo) +
            o.g1_msm(c["A"][64 * lo:64 * hi], w[32 * lo:32 * hi], hi - lo) +
            o.g1_msm(c["B1"][64 * lo:64 * hi], w[32 * lo:32 * hi], hi - lo) +
            o.g2_msm(c["B2"][128 * lo:128 * hi], w[32 * lo:32 * hi], hi - lo) +
            o.g1_msm(c["C"][64 * (clo - skip):64 * (chi - skip)], w[32 * clo:32 * chi], chi - clo))
    folded, proof = bdist.finish_proof(part, c["vk"], bytes.fromhex(C["r"]), bytes.fromhex(C["s"]))
    q.put((rank, proof.hex()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_proof_over_gloo_matches_golden(world):
    """N > 1 host path on CPU: per-rank partial MSMs -> all_gather (gloo) -> fold -> finalize == unsharded proof."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(pr == C["proof"] for _, pr in res)


def _exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from rapidsnark_old_b200 import dist as bdist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    owners = bdist.poly_owners(world)
    # every rank fills only the polynomials it owns (value = 10 * poly + owner), the rest is garbage
    n = 24                                   # "domain" of 24 elements of 32 bytes: 24 splits evenly and unevenly
    polys = [torch.full((n * 32,), 10 * i + o if o == rank else 255, dtype=torch.uint8) for i, o in enumerate(owners)]
    bdist.exchange_polys(polys)
    lo, hi = bdist.plan_range(n, rank, world)
    mine = [p[lo * 32:hi * 32] for p in polys]            # the slice this rank combines
    q.put((rank, bdist.poly_mask(rank, world), [int(p[0]) for p in mine], [bool((p == p[0]).all()) for p in mine]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 5])
def test_h_polynomial_exchange_over_gloo(world):
    """N > 1 host logic of the H-pipeline split: ownership masks cover a, b, c exactly once and the point-to-point
    exchange leaves every rank with ITS slice of all three (dist.exchange_polys; NCCL on the GPU box, gloo here)."""
    import torch.multiprocessing as mp
    from rapidsnark_old_b200 import dist as bdist
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    owners = bdist.poly_owners(world)
    assert sum(m for _, m, _, _ in res) == 7 and all((res[o][1] >> i) & 1 for i, o in enumerate(owners))
    for _, _, vals, uniform in res:
        assert vals == [10 * i + o for i, o in enumerate(owners)] and all(uniform)


# ----------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def gctx():
    c = b200.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_primitives_match_golden(gctx):
    m = G["msm_g1"]
    got = gctx.msm_g1(bytes.fromhex(m["bases"]), bytes.fromhex(m["scalars"]), m["n"])
    assert b200.host_g1_to_affine(got).hex() == m["affine"]
    m = G["msm_g2"]
    got = gctx.msm_g2(bytes.fromhex(m["bases"]), bytes.fromhex(m["scalars"]), m["n"])
    assert b200.host_g2_to_affine(got).hex() == m["affine"]
    t = G["ntt"]
    assert gctx.ntt(bytes.fromhex(t["in"])).hex() == t["fft"]
    assert gctx.ntt(bytes.fromhex(t["in"]), inverse=True).hex() == t["ifft"]


@pytest.mark.gpu
def test_gpu_prove_matches_golden(gctx):
    c = _circuit()
    zk = gctx.zkey_upload(c["n_vars"], c["n_public"], c["domain"], c["n_coefs"], c["coefs"], c["A"], c["B1"], c["B2"],
                          c["C"], c["H"])
    assert zk.h_scalars(c["wtns"]).hex() == C["h_scalars"]
    msms = zk.prove_msms(c["wtns"])
    aff = [b200.host_g1_to_affine(msms[0:128]), b200.host_g1_to_affine(msms[128:256]),
           b200.host_g1_to_affine(msms[256:384]), b200.host_g2_to_affine(msms[384:640]),
           b200.host_g1_to_affine(msms[640:768])]
    assert [a.hex() for a in aff] == C["msms_affine"]
    proof = b200.groth16_finalize(msms, c["vk"], bytes.fromhex(C["r"]), bytes.fromhex(C["s"]))
    assert proof.hex() == C["proof"]
    zk.free()


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", ["1", "2", "3", "2-replicated"])
def test_cli_output_is_byte_identical_to_reference_cli(tmp_path, gpus):
    """build/prover <zkey> <wtns> <proof.json> <public.json> vs the reference CLI's files for the same r, s.
    B200_GPUS=N runs N point-range shards from one process: stage 1 per GPU, device-to-device exchange of the
    transformed polynomials (b200_exchange_polys), stage 2 - on one physical GPU if fewer are present
    (B200_DEVICE_STRIDE=0)."""
    replicate = gpus.endswith("-replicated")
    gpus = gpus.split("-")[0]
    zk, wt = tmp_path / "c.zkey", tmp_path / "w.wtns"
    zk.write_bytes(bytes.fromhex(C["zkey"]))
    wt.write_bytes(bytes.fromhex(C["wtns"]))
    env = dict(os.environ, B200_R=bytes.fromhex(C["r"])[::-1].hex(), B200_S=bytes.fromhex(C["s"])[::-1].hex())
    if gpus != "1":
        import torch
        if torch.cuda.device_count() < int(gpus):
            env["B200_DEVICE_STRIDE"] = "0"
        env["B200_GPUS"] = gpus
        if replicate:
            env["B200_REPLICATE_H"] = "1"
    r = subprocess.run([PROVER, str(zk), str(wt), str(tmp_path / "proof.json"), str(tmp_path / "public.json")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "proof.json").read_text() == C["proof_json"]
    assert (tmp_path / "public.json").read_text() == C["public_json"]


SEAM_CLI = os.path.join(ROOT, "oracle", "_ref", "ref_prover_b200seam")


def _run_seam_cli(tmp_path):
    zk, wt = tmp_path / "c.zkey", tmp_path / "w.wtns"
    zk.write_bytes(bytes.fromhex(C["zkey"]))
    wt.write_bytes(bytes.fromhex(C["wtns"]))
    env = dict(os.environ, B200_R=bytes.fromhex(C["r"])[::-1].hex(), B200_S=bytes.fromhex(C["s"])[::-1].hex())
    return subprocess.run([SEAM_CLI, str(zk), str(wt), str(tmp_path / "proof.json"), str(tmp_path / "public.json")],
                          capture_output=True, text=True, env=env, cwd=tmp_path)


@pytest.mark.skipif(not os.path.exists(SEAM_CLI), reason="oracle/_ref not built (no /root/reference where build() ran)")
def test_reference_cli_builds_against_the_engine_headers_and_needs_a_gpu(tmp_path):
    """INTEGRATION.md, inner seam: the reference's src/main_prover.cpp, groth16.cpp (through groth16.hpp), logger,
    binfile / zkey / wtns readers - all unmodified - compiled against rapidsnark_old_b200/src/engine/ instead of
    depends/ffiasm/c (oracle/Makefile: _ref/ref_prover_b200seam).  Without a GPU it parses the files, builds a, b, c on
    the host and then fails at the first FFT / MSM call: there is no CPU path behind the seam."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present: the GPU test below runs it for real")
    except ImportError:
        pass
    r = _run_seam_cli(tmp_path)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr or "no CPU path" in r.stderr, r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SEAM_CLI), reason="oracle/_ref not built (no /root/reference where build() ran)")
def test_reference_cli_on_the_gpu_seam_is_byte_identical(tmp_path):
    """The same binary on the GPU box: the reference's own prove() flow with Curve::multiMulByScalar and FFT::fft /
    ifft served by libb200snark.so writes the reference CLI's proof.json / public.json byte for byte (same r, s)."""
    r = _run_seam_cli(tmp_path)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "proof.json").read_text() == C["proof_json"]
    assert (tmp_path / "public.json").read_text() == C["public_json"]


@pytest.mark.gpu
def test_fullprover_status_machine(tmp_path):
    """FullProver (src/fullprover.cpp:21-240): ready -> busy -> success, witness from ./build/<circuit> via popen,
    status JSON {"proof": "<json text>", "pubData": "<json text>", "status": "success"}; the proof verifies."""
    import stat
    import pairing
    from test_verifier import _proof_from_json, _vk_and_public
    (tmp_path / "build").mkdir()
    zk, wt = tmp_path / "mycircuit.zkey", tmp_path / "w.wtns"
    zk.write_bytes(bytes.fromhex(C["zkey"]))
    wt.write_bytes(bytes.fromhex(C["wtns"]))
    gen = tmp_path / "build" / "mycircuit"          # stand-in for the circom witness calculator binary
    gen.write_text("#!/bin/sh\ncp %s \"$2\"\necho witness done\n" % wt)
    gen.chmod(gen.stat().st_mode | stat.S_IEXEC)
    inp = tmp_path / "input.json"
    inp.write_text('{"a": 1}')
    r = subprocess.run([os.path.join(ROOT, "build", "fullprover_demo"), str(zk), "--", "mycircuit", str(inp)],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines[0] == '{"status":"ready"}'
    st = json.loads(lines[-1])
    assert st["status"] == "success"
    assert json.loads(st["pubData"]) == json.loads(C["public_json"])
    c, vk, public = _vk_and_public()
    assert pairing.groth16_verify(vk, _proof_from_json(st["proof"]), public)
    # unknown circuit -> failed with an error text, like the reference's catch(std::runtime_error)
    r = subprocess.run([os.path.join(ROOT, "build", "fullprover_demo"), str(zk), "--", "nosuch", str(inp)],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and '"status":"failed"' in r.stdout


@pytest.mark.gpu
def test_proof_server_routes(tmp_path):
    """build/proverServer <port> <zkey>: the reference's REST routes (src/main_proofserver.cpp:36-40, proverapi.cpp:9-41)
    - GET /status, POST /input/:circuit, POST /cancel, /start, /stop - and its status documents; the proof verifies."""
    import socket
    import stat
    import time
    import urllib.request
    import urllib.error
    from rapidsnark_old_b200 import verify
    sys_path_tools = os.path.join(ROOT, "tools")
    import sys
    sys.path.insert(0, sys_path_tools)
    import export_vkey
    (tmp_path / "build").mkdir()
    zk, wt = tmp_path / "mycircuit.zkey", tmp_path / "w.wtns"
    zk.write_bytes(bytes.fromhex(C["zkey"]))
    wt.write_bytes(bytes.fromhex(C["wtns"]))
    gen = tmp_path / "build" / "mycircuit"          # stand-in for the circom witness calculator binary
    gen.write_text("#!/bin/sh\ncp %s \"$2\"\necho witness done\n" % wt)
    gen.chmod(gen.stat().st_mode | stat.S_IEXEC)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    srv = subprocess.Popen([os.path.join(ROOT, "build", "proverServer"), str(port), str(zk)], cwd=tmp_path,
                           env=dict(os.environ, B200_SERVER_ALLOW_QUIT="1"), stderr=subprocess.PIPE, text=True)
    base = "http://127.0.0.1:%d" % port

    def call(method, path, body=None):
        req = urllib.request.Request(base + path, data=body, method=method)
        with urllib.request.urlopen(req, timeout=30) as r:
            return r.status, r.read().decode()
    try:
        for _ in range(600):                         # zkey upload + CUDA context creation
            try:
                code, text = call("GET", "/status")
                break
            except (urllib.error.URLError, ConnectionError):
                assert srv.poll() is None, srv.stderr.read()
                time.sleep(0.1)
        assert code == 200 and json.loads(text) == {"status": "ready"}
        assert call("POST", "/start")[0] == 200 and call("POST", "/stop")[0] == 200
        assert call("POST", "/input/mycircuit", b'{"a": 1}')[0] == 200
        for _ in range(600):
            st = json.loads(call("GET", "/status")[1])
            if st["status"] != "busy":
                break
            time.sleep(0.05)
        assert st["status"] == "success", st
        assert json.loads(st["pubData"]) == json.loads(C["public_json"])
        assert verify.verify(export_vkey.export(zk.read_bytes()), json.loads(st["pubData"]), json.loads(st["proof"]))
        assert call("POST", "/input/mycircuit", b'{"a": ')[0] == 200          # malformed request body
        for _ in range(600):
            st = json.loads(call("GET", "/status")[1])
            if st["status"] != "busy":
                break
            time.sleep(0.05)
        assert st["status"] == "failed" and "parse error" in st["error"]
        assert call("POST", "/cancel")[0] == 200
        with pytest.raises(urllib.error.HTTPError):
            call("GET", "/nosuch")
        call("POST", "/quit")
        srv.wait(timeout=20)
    finally:
        if srv.poll() is None:
            srv.kill()
