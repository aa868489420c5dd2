#!/usr/bin/env python3
"""BASELINE.json configs #1 (single-blob latency), #2, #4 and #5 at their full sizes, end to end through the
host-buffer C ABI on one GPU, each with a parity check of >= 64 items against the CPU oracle (oracle/, the checker -- never
the thing measured) and the oracle's own rate beside it.  bench.py imports these functions and puts their results into its
JSON line under "configs"; run stand-alone it prints one JSON line per config:

  python tools/bench_configs.py [--reps 3] [--parity 64] [--only latency,eip4844,recover,verify,pageable]"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_BLOB, CELL, NCELLS = 131072, 2048, 128
# multiply-adds on the fmaheavy pipe: XYZZ mixed addition with one fused reduction
OP_XYZZ_MADD = 8 * 300 + 2 * 234 - 156


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def adds_per_scalar(w):
    """table additions per 255-bit scalar at window width w with the merged top window (csrc/msm_table.cuh)"""
    nw = 255 // w + 1
    rtop = (1 << (255 - w * (nw - 1))) + 1
    mg = next((m for m in (4, 2) if rtop ** m - 1 <= 1 << (w - 1)), 1)
    return nw - 1 + 1.0 / mg


def best(fn, reps):
    ts = []
    r = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def _pinned(nbytes, src=None):
    import torch
    t = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
    if src is not None:
        t[:len(src)] = torch.frombuffer(bytearray(src), dtype=torch.uint8)
    return t


def _ok(res):
    assert res.status == 0, C.cast(res.error_msg, C.c_char_p).value


def _P(t):
    return C.c_void_p(t.data_ptr())


def _pool_map(fn, items):
    """oracle calls release the GIL (ctypes), so a thread pool uses all host cores"""
    with ThreadPoolExecutor(max_workers=host_threads()) as ex:
        return list(ex.map(fn, items))


def _env(ctx, pkg):
    import importlib
    from oracle import cref
    cref.build()
    return pkg.load_library(), C.c_void_p(ctx.handle), importlib.import_module("eth_kzg_b200.synthetic"), cref


def config2(ctx, pkg, reps=3, parity=64, imad_peak=None):
    """#2 blob_to_kzg_commitment + compute_blob_kzg_proof, batch of 1024 blobs"""
    lib, H, syn, cref = _env(ctx, pkg)
    n = 1024
    blobs = syn.blobs(n)
    h_blobs, h_comm, h_proof, h_st = _pinned(len(blobs), blobs), _pinned(n * 48), _pinned(n * 48), _pinned(n)
    commit = lambda: _ok(lib.eth_kzg_b200_blob_to_kzg_commitment_batch(H, C.c_uint64(n), _P(h_blobs), _P(h_comm), _P(h_st)))
    prove = lambda: _ok(lib.eth_kzg_b200_compute_blob_kzg_proof_batch(H, C.c_uint64(n), _P(h_blobs), _P(h_comm), _P(h_proof), _P(h_st)))
    commit(); prove()
    t1, _ = best(commit, reps)
    t2, _ = best(prove, reps)
    comms, proofs = bytes(h_comm.numpy()), bytes(h_proof.numpy())
    k = min(parity, n)

    def one(i):
        b = blobs[i * BYTES_PER_BLOB:(i + 1) * BYTES_PER_BLOB]
        c = cref.blob_to_kzg_commitment(b)
        return c == comms[48 * i:48 * i + 48] and cref.compute_blob_kzg_proof(b, c) == proofs[48 * i:48 * i + 48]
    t0 = time.perf_counter()
    good = _pool_map(one, range(k))
    tc = time.perf_counter() - t0
    assert all(good), "config #2: GPU and oracle disagree on blobs %r" % [i for i, g in enumerate(good) if not g]
    imad = 4096 * adds_per_scalar(ctx.srs_window) * OP_XYZZ_MADD   # one 4096-point fixed-base MSM over the SRS tables, per blob
    d = {"workload": "blob_to_kzg_commitment + compute_blob_kzg_proof, batch of 1024 synthetic blobs (host buffers through the C ABI)",
         "metric": "blobs/s", "value": n / (t1 + t2), "ms": 1e3 * (t1 + t2), "commit_ms": 1e3 * t1, "blob_proof_ms": 1e3 * t2,
         "dominant_kernel": "k_fk20_msm over the SRS tables (w=%d: 4096 x %.2f additions x %d multiply-adds per blob and per MSM)" % (ctx.srs_window, adds_per_scalar(ctx.srs_window), OP_XYZZ_MADD),
         "cpu_port": {"value": k / tc, "unit": "blobs/s", "cores": host_threads(), "sample": "%d blobs (commitment + blob proof each)" % k},
         "parity_checked": k}
    if imad_peak:
        d["imad_frac_of_call"] = {"commit": n * imad / t1 / imad_peak, "blob_proof": n * imad / t2 / imad_peak,
                                  "note": "MSM multiply-adds / whole host-to-host call time / measured IMAD.WIDE peak (copies, SHA-256 challenge and quotient included in the time)"}
    return d


def config4(ctx, pkg, reps=3, parity=64, imad_peak=None, fk20_imad_per_blob=None):
    """#4 recover_cells_and_kzg_proofs with 50 % of the 128 cells erased, batch of 256 blobs"""
    lib, H, syn, cref = _env(ctx, pkg)
    n = 256
    blobs = syn.blobs(n)
    cells_flat, proofs_flat, st = ctx.compute_cells_and_kzg_proofs_batch(blobs, n)
    assert not any(st)
    patterns = {"first_half_missing": list(range(64, 128)), "second_half_missing": list(range(0, 64)), "every_other": list(range(0, 128, 2))}
    res = {}
    h_oc, h_op, h_st = _pinned(n * NCELLS * CELL), _pinned(n * NCELLS * 48), _pinned(n)
    for name, keep in patterns.items():
        counts = (C.c_uint64 * n)(*([len(keep)] * n))
        idx = (C.c_uint64 * (n * len(keep)))(*(keep * n))
        h_in = _pinned(n * len(keep) * CELL, b"".join(cells_flat[(b * NCELLS + i) * CELL:(b * NCELLS + i + 1) * CELL] for b in range(n) for i in keep))
        call = lambda: _ok(lib.eth_kzg_b200_recover_cells_and_kzg_proofs_batch(H, C.c_uint64(n), counts, idx, _P(h_in), _P(h_oc), _P(h_op), _P(h_st)))
        call()
        t, _ = best(call, reps)
        assert bytes(h_oc.numpy()) == cells_flat and bytes(h_op.numpy()) == proofs_flat, "recovery round trip failed: " + name
        res[name + "_ms"] = 1e3 * t
    k = min(parity, n)
    keep = patterns["every_other"]

    def one(b):
        oc, op = cref.recover_cells_and_kzg_proofs(keep, [cells_flat[(b * NCELLS + i) * CELL:(b * NCELLS + i + 1) * CELL] for i in keep])
        return b"".join(oc) == cells_flat[b * NCELLS * CELL:(b + 1) * NCELLS * CELL] and b"".join(op) == proofs_flat[b * NCELLS * 48:(b + 1) * NCELLS * 48]
    t0 = time.perf_counter()
    good = _pool_map(one, range(k))
    tc = time.perf_counter() - t0
    assert all(good), "config #4: GPU and oracle disagree"
    worst = max(res.values())
    d = dict({"workload": "recover_cells_and_kzg_proofs, 64 of 128 cells erased, batch of 256 synthetic blobs (host buffers through the C ABI); value = slowest of three erasure patterns",
              "metric": "blobs/s", "value": n / (worst / 1e3), "ms": worst,
              "dominant_kernel": "k_fk20_msm + k_fk20_g1_ntts (the proofs of the recovered polynomial are a full FK20 pass; the recovery transforms themselves are < 2 ms)",
              "cpu_port": {"value": k / tc, "unit": "blobs/s", "cores": host_threads(), "sample": "%d blobs, every other cell missing" % k},
              "parity_checked": k, "round_trip_checked": n}, **res)
    if imad_peak and fk20_imad_per_blob:
        d["imad_frac_of_call"] = n * fk20_imad_per_blob / (worst / 1e3) / imad_peak
    return d


def config5(ctx, pkg, reps=3, parity=64):
    """#5 verify_cell_kzg_proof_batch over 128 blobs x 128 cells, and the reference's own bench shape (1 blob x 128 cells)"""
    lib, H, syn, cref = _env(ctx, pkg)
    n = 128
    blobs = syn.blobs(n)
    cells_flat, proofs_flat, st = ctx.compute_cells_and_kzg_proofs_batch(blobs, n)
    comms, st2 = ctx.blob_to_kzg_commitment_batch(blobs, n)
    assert not any(st) and not any(st2)
    N = n * NCELLS
    # the reference ABI takes arrays of pointers to individual items (bindings/c/src/lib.rs:309): build them once
    b_comm, b_cells, b_proofs = C.create_string_buffer(comms, n * 48), C.create_string_buffer(cells_flat, N * CELL), C.create_string_buffer(proofs_flat, N * 48)
    a0, a1, a2 = C.addressof(b_comm), C.addressof(b_cells), C.addressof(b_proofs)
    pc = (C.c_void_p * N)(*[a0 + 48 * (k // NCELLS) for k in range(N)])
    pl = (C.c_void_p * N)(*[a1 + CELL * k for k in range(N)])
    pp = (C.c_void_p * N)(*[a2 + 48 * k for k in range(N)])
    idx = (C.c_uint64 * N)(*[k % NCELLS for k in range(N)])
    flag = C.c_bool(False)

    def call(count=N):
        _ok(lib.eth_kzg_verify_cell_kzg_proof_batch(H, C.c_uint64(count), pc, C.c_uint64(count), idx, C.c_uint64(count), pl, C.c_uint64(count), pp, C.byref(flag)))
        return bool(flag.value)
    call()
    t, good = best(call, reps)
    assert good is True
    t1, good1 = best(lambda: call(NCELLS), max(reps, 5))      # one blob's 128 cells (crates/eip7594/benches/benchmark-mt.rs:77-101)
    assert good1 is True
    bad_cell = 777
    orig = b_cells[bad_cell * CELL + CELL - 1]
    b_cells[bad_cell * CELL + CELL - 1] = bytes([cells_flat[bad_cell * CELL + CELL - 1] ^ 1])
    tneg, okneg = best(call, 1)
    assert okneg is False, "a corrupted cell was accepted"
    b_cells[bad_cell * CELL + CELL - 1] = orig
    k = min(parity, n)

    def one(b):
        """the oracle on blob b's 128 cells, intact (must accept) -- and with the corrupted cell where it lies (must reject)"""
        cl = [cells_flat[(b * NCELLS + i) * CELL:(b * NCELLS + i + 1) * CELL] for i in range(NCELLS)]
        pf = [proofs_flat[(b * NCELLS + i) * 48:(b * NCELLS + i + 1) * 48] for i in range(NCELLS)]
        cm = [comms[48 * b:48 * b + 48]] * NCELLS
        okc = cref.verify_cell_kzg_proof_batch(cm, list(range(NCELLS)), cl, pf)
        if b == bad_cell // NCELLS:
            i = bad_cell % NCELLS
            cl[i] = cl[i][:-1] + bytes([cl[i][-1] ^ 1])
            okc = okc and cref.verify_cell_kzg_proof_batch(cm, list(range(NCELLS)), cl, pf) is False
        return okc is True
    t0 = time.perf_counter()
    good = _pool_map(one, range(k))
    tc = time.perf_counter() - t0
    assert all(good), "config #5: the oracle rejects cells the GPU accepted (or accepts the corrupted one)"
    t0 = time.perf_counter()
    one(1)
    tc1 = time.perf_counter() - t0
    return {"workload": "verify_cell_kzg_proof_batch, 128 blobs x 128 cells in one call (the reference's pointer-array C ABI)", "metric": "cells/s",
            "value": N / t, "ms": 1e3 * t, "negative_case_ms": 1e3 * tneg,
            "one_blob_128_cells_ms": 1e3 * t1, "cpu_port_one_blob_128_cells_ms": 1e3 * tc1,
            "dominant_cost": "the consensus-mandated single SHA-256 chain over 34.6 MB of cells (host, SHA-NI, ~21 ms) -- the floor of this call; the device side "
                             "(decompression, subgroup checks, interpolation, the two RLC sums) runs under it",
            "cpu_port": {"value": k * NCELLS / tc, "unit": "cells/s", "cores": host_threads(), "sample": "%d calls of 128 cells" % k},
            "parity_checked": k * NCELLS}


def latency(ctx, pkg, reps=5):
    """config #1: ONE blob (and 32 concurrent single-blob callers) through the reference's own symbol
    eth_kzg_compute_cells_and_kzg_proofs, beside the CPU oracle's single-blob time"""
    lib, H, syn, cref = _env(ctx, pkg)
    T = 32
    blobs = [syn.blob(9000 + i) for i in range(T)]
    bufs = []
    for _ in range(T):
        cells = [C.create_string_buffer(CELL) for _ in range(NCELLS)]
        proofs = [C.create_string_buffer(48) for _ in range(NCELLS)]
        bufs.append((cells, proofs, (C.c_void_p * NCELLS)(*[C.addressof(b) for b in cells]), (C.c_void_p * NCELLS)(*[C.addressof(b) for b in proofs])))

    def call(i):
        _ok(lib.eth_kzg_compute_cells_and_kzg_proofs(H, blobs[i], bufs[i][2], bufs[i][3]))
    call(0)
    one = sorted(best(lambda: call(0), 1)[0] for _ in range(reps))

    def wave():
        barrier = threading.Barrier(T + 1)
        th = [threading.Thread(target=lambda i=i: (barrier.wait(), call(i))) for i in range(T)]
        for t in th:
            t.start()
        barrier.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        return time.perf_counter() - t0
    wave()
    many = sorted(wave() for _ in range(reps))
    oc, op = cref.compute_cells_and_kzg_proofs(blobs[0])
    t0 = time.perf_counter()
    cref.compute_cells_and_kzg_proofs(blobs[0])
    tc = time.perf_counter() - t0
    got_c = [b.raw for b in bufs[0][0]]
    got_p = [b.raw for b in bufs[0][1]]
    assert got_c == list(oc) and got_p == list(op), "single-blob call differs from the oracle"

    # the other per-item symbols of the reference's ABI, one call each from one thread (median of `reps`), GPU | CPU oracle port
    def med(fn, n=reps):
        fn()
        ts = sorted(best(fn, 1)[0] for _ in range(n))
        return 1e3 * ts[len(ts) // 2]

    def both(gpu_fn, cpu_fn):
        g = gpu_fn()
        t0 = time.perf_counter()
        c = cpu_fn()
        tcpu = 1e3 * (time.perf_counter() - t0)
        assert g == c, "GPU and oracle disagree"
        return {"gpu_ms": med(gpu_fn), "cpu_port_ms_1thread": tcpu}
    b0 = blobs[0]
    cm = ctx.blob_to_kzg_commitment(b0)
    z = (12345).to_bytes(32, "big")
    pf_blob = ctx.compute_blob_kzg_proof(b0, cm)
    pf_z, y = ctx.compute_kzg_proof(b0, z)
    keep = list(range(0, NCELLS, 2))
    per_item = {
        "blob_to_kzg_commitment": both(lambda: ctx.blob_to_kzg_commitment(b0), lambda: cref.blob_to_kzg_commitment(b0)),
        "compute_blob_kzg_proof": both(lambda: ctx.compute_blob_kzg_proof(b0, cm), lambda: cref.compute_blob_kzg_proof(b0, cm)),
        "compute_kzg_proof": both(lambda: tuple(ctx.compute_kzg_proof(b0, z)), lambda: tuple(cref.compute_kzg_proof(b0, z))),
        "verify_kzg_proof": both(lambda: ctx.verify_kzg_proof(cm, z, y, pf_z), lambda: cref.verify_kzg_proof(cm, z, y, pf_z)),
        "verify_blob_kzg_proof": both(lambda: ctx.verify_blob_kzg_proof(b0, cm, pf_blob), lambda: cref.verify_blob_kzg_proof(b0, cm, pf_blob)),
        "recover_cells_and_kzg_proofs_64_of_128": both(lambda: ctx.recover_cells_and_kzg_proofs(keep, [got_c[i] for i in keep]),
                                                       lambda: tuple(list(x) for x in cref.recover_cells_and_kzg_proofs(keep, [got_c[i] for i in keep]))),
        "verify_cell_kzg_proof_batch_128_cells": both(lambda: ctx.verify_cell_kzg_proof_batch([cm] * NCELLS, list(range(NCELLS)), got_c, got_p),
                                                      lambda: cref.verify_cell_kzg_proof_batch([cm] * NCELLS, list(range(NCELLS)), got_c, got_p)),
    }
    # a block's worth of blobs in ONE batch call (preallocated host buffers): FK20 in latency mode (cooperative G1 NTTs)
    nb = 8
    src = bytearray(b"".join(blobs[:nb]))
    bc, bp, bs = bytearray(nb * NCELLS * CELL), bytearray(nb * NCELLS * 48), bytearray(nb)
    cb, cc_, cp_, cs_ = [(C.c_char * len(x)).from_buffer(x) for x in (src, bc, bp, bs)]
    batch8 = lambda: _ok(lib.eth_kzg_b200_compute_cells_and_kzg_proofs_batch(H, C.c_uint64(nb), cb, cc_, cp_, cs_))
    batch8()
    eight = sorted(best(batch8, 1)[0] for _ in range(reps))
    assert bytes(bc[:NCELLS * CELL]) == b"".join(oc) and bytes(bp[:NCELLS * 48]) == b"".join(op), "8-blob batch differs from the oracle"
    t0 = time.perf_counter()
    cref.compute_cells_and_kzg_proofs_batch(bytes(src), nb)
    tc8 = time.perf_counter() - t0
    return {"workload": "compute_cells_and_kzg_proofs for 1 blob through eth_kzg_compute_cells_and_kzg_proofs (BASELINE config #1)",
            "latency_8blob_batch_ms": 1e3 * eight[len(eight) // 2], "cpu_port_8blob_batch_ms": 1e3 * tc8, "cpu_port_8blob_threads": cref.num_threads(),
            "latency_8blob_note": "8 blobs (a block's worth) in one eth_kzg_b200_compute_cells_and_kzg_proofs_batch call, host buffers",
            "latency_1blob_ms": 1e3 * one[len(one) // 2], "latency_1blob_ms_min": 1e3 * one[0],
            "latency_32blob_ms": 1e3 * many[len(many) // 2], "latency_32blob_note": "32 host threads, one single-blob call each, started together; until the last returns",
            "cpu_port_1blob_ms": 1e3 * tc, "cpu_port_threads": 1, "parity_checked": 1,
            "per_item_symbols_ms": per_item,
            "per_item_note": "one call from one thread through the Python ctypes stub (list marshalling of 128 cells included), median of %d" % reps}


def pageable(ctx, pkg, reps=3):
    """config #3 through the batch entry point with PAGEABLE caller buffers (what a binding that does not pin its memory hands
    over): the library stages through its own pinned memory under the kernels"""
    lib, H, syn, cref = _env(ctx, pkg)
    n = 1024
    src = bytearray(syn.blobs(n))
    cells, proofs, status = bytearray(n * 262144), bytearray(n * 6144), bytearray(n)
    cb = (C.c_char * len(src)).from_buffer(src)
    cc = (C.c_char * len(cells)).from_buffer(cells)
    cp = (C.c_char * len(proofs)).from_buffer(proofs)
    cs = (C.c_char * len(status)).from_buffer(status)
    call = lambda: _ok(lib.eth_kzg_b200_compute_cells_and_kzg_proofs_batch(H, C.c_uint64(n), cb, cc, cp, cs))
    call()
    t, _ = best(call, reps)
    assert bytes(cells[:131072]) == bytes(src[:131072])
    return {"workload": "compute_cells_and_kzg_proofs, batch of 1024 blobs, PAGEABLE host buffers through the C ABI", "metric": "blobs/s", "value": n / t, "ms": 1e3 * t}


def abi_load(threads=1024, calls=16, mode="compute", timeout=600, env=None):
    """tools/abi_load.c (native pthreads, no Python): T callers of the reference's per-item symbol on one shared context.
    Creates its own context, so the caller must have released the device memory of any other."""
    exe = os.path.join(ROOT, "rust-eth-kzg_b200", "lib", "abi_load")
    if not os.path.exists(exe):
        return {"failed": "lib/abi_load not built"}
    import subprocess
    e = dict(os.environ)
    e.update(env or {})
    try:
        out = subprocess.run([exe, "--threads", str(threads), "--calls", str(calls), "--mode", mode], capture_output=True, text=True, timeout=timeout, env=e)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if not line:
            return {"failed": "no output (rc %d): %s" % (out.returncode, out.stderr[-300:])}
        return json.loads(line[-1])
    except Exception as ex:
        return {"failed": repr(ex)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--parity", type=int, default=64)
    ap.add_argument("--only", default="latency,eip4844,recover,verify,pageable")
    args = ap.parse_args()
    only = set(args.only.split(","))
    import __graft_entry__
    pkg = __graft_entry__.load_package()
    ctx = pkg.DASContext(use_precomp=True)
    peak = ctx.probe_imad_wide() or None

    def emit(d):
        d.update({"n_gpus": 1, "data": "synthetic", "fk20_window_bits": ctx.window, "srs_window_bits": ctx.srs_window})
        print(json.dumps(d), flush=True)
    if "latency" in only:
        emit(latency(ctx, pkg))
    if "eip4844" in only:
        emit(config2(ctx, pkg, args.reps, args.parity, peak))
    if "recover" in only:
        emit(config4(ctx, pkg, args.reps, args.parity, peak))
    if "verify" in only:
        emit(config5(ctx, pkg, args.reps, args.parity))
    if "pageable" in only:
        emit(pageable(ctx, pkg, args.reps))
    ctx.close()
    if "abi" in only:
        print(json.dumps(abi_load()), flush=True)


if __name__ == "__main__":
    main()
