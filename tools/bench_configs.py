#!/usr/bin/env python3
"""Times BASELINE.json configs #2, #4 and #5 end to end through the host-buffer C ABI on one GPU and,
on a bounded sample, through the CPU oracle port.  One JSON line per config; the headline config #3 is bench.py.

  python tools/bench_configs.py [--reps 3] [--cpu-sample 8] [--only eip4844,recover,verify]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

BYTES_PER_BLOB, CELL, NCELLS = 131072, 2048, 128


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=8)
    ap.add_argument("--only", default="eip4844,recover,verify,callers,pageable")
    args = ap.parse_args()
    only = set(args.only.split(","))
    pkg = __graft_entry__.load_package()
    import importlib
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    from oracle import cref
    cref.build()
    ctx = pkg.DASContext(use_precomp=True)
    cores = cref.num_threads()
    out = []

    def emit(d):
        d.update({"n_gpus": 1, "data": "synthetic", "cpu_cores": cores, "fk20_window_bits": ctx.window})
        print(json.dumps(d), flush=True)
        out.append(d)

    import ctypes as C
    import torch
    lib = pkg.load_library()
    H = C.c_void_p(ctx.handle)

    def pinned(nbytes, src=None):
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
        if src is not None:
            t[:len(src)] = torch.frombuffer(bytearray(src), dtype=torch.uint8)
        return t

    def ok(res):
        assert res.status == 0, C.cast(res.error_msg, C.c_char_p).value

    if "eip4844" in only:
        n = 1024
        blobs = syn.blobs(n)
        h_blobs, h_comm, h_proof, h_st = pinned(len(blobs), blobs), pinned(n * 48), pinned(n * 48), pinned(n)
        P = lambda t: C.c_void_p(t.data_ptr())
        commit = lambda: ok(lib.eth_kzg_b200_blob_to_kzg_commitment_batch(H, C.c_uint64(n), P(h_blobs), P(h_comm), P(h_st)))
        prove = lambda: ok(lib.eth_kzg_b200_compute_blob_kzg_proof_batch(H, C.c_uint64(n), P(h_blobs), P(h_comm), P(h_proof), P(h_st)))
        commit(); prove()  # warm-up
        t1, _ = best(commit, args.reps)
        t2, _ = best(prove, args.reps)
        comms, proofs = bytes(h_comm.numpy()), bytes(h_proof.numpy())
        k = args.cpu_sample
        t0 = time.perf_counter()
        for i in range(k):
            b = blobs[i * BYTES_PER_BLOB:(i + 1) * BYTES_PER_BLOB]
            c = cref.blob_to_kzg_commitment(b)
            p = cref.compute_blob_kzg_proof(b, c)
            assert c == comms[48 * i:48 * i + 48] and p == proofs[48 * i:48 * i + 48], "GPU and oracle disagree on blob %d" % i
        tc = time.perf_counter() - t0
        emit({"config": "#2 blob_to_kzg_commitment + compute_blob_kzg_proof, batch of 1024 blobs (host buffers through the C ABI)", "metric": "blobs/s",
              "value": n / (t1 + t2), "commit_ms": 1e3 * t1, "blob_proof_ms": 1e3 * t2, "cpu_port_blobs_per_s_1thread": k / tc, "parity_checked_blobs": k})

    if "recover" in only:
        n = 256
        blobs = syn.blobs(n)
        cells_flat, proofs_flat, st = ctx.compute_cells_and_kzg_proofs_batch(blobs, n)
        assert not any(st)
        patterns = {"first_half_missing": list(range(64, 128)), "second_half_missing": list(range(0, 64)), "every_other": list(range(0, 128, 2))}
        res = {}
        h_oc, h_op, h_st = pinned(n * NCELLS * CELL), pinned(n * NCELLS * 48), pinned(n)
        P = lambda t: C.c_void_p(t.data_ptr())
        for name, keep in patterns.items():
            counts = (C.c_uint64 * n)(*([len(keep)] * n))
            idx = (C.c_uint64 * (n * len(keep)))(*(keep * n))
            h_in = pinned(n * len(keep) * CELL, b"".join(cells_flat[(b * NCELLS + i) * CELL:(b * NCELLS + i + 1) * CELL] for b in range(n) for i in keep))
            call = lambda: ok(lib.eth_kzg_b200_recover_cells_and_kzg_proofs_batch(H, C.c_uint64(n), counts, idx, P(h_in), P(h_oc), P(h_op), P(h_st)))
            call()
            t, _ = best(call, args.reps)
            assert bytes(h_oc.numpy()) == cells_flat and bytes(h_op.numpy()) == proofs_flat, "recovery round trip failed: " + name
            res[name + "_ms"] = 1e3 * t
        k = max(1, args.cpu_sample // 4)
        keep = patterns["first_half_missing"]
        t0 = time.perf_counter()
        for b in range(k):
            oc, op = cref.recover_cells_and_kzg_proofs(keep, [cells_flat[(b * NCELLS + i) * CELL:(b * NCELLS + i + 1) * CELL] for i in keep])
            assert b"".join(oc) == cells_flat[b * NCELLS * CELL:(b + 1) * NCELLS * CELL] and b"".join(op) == proofs_flat[b * NCELLS * 48:(b + 1) * NCELLS * 48]
        tc = time.perf_counter() - t0
        worst = max(res.values())
        emit(dict({"config": "#4 recover_cells_and_kzg_proofs, 64 of 128 cells erased, batch of 256 blobs (host buffers through the C ABI)", "metric": "blobs/s",
                   "value": n / (worst / 1e3), "cpu_port_blobs_per_s_1thread": k / tc, "parity_checked_blobs": k, "round_trip_checked_blobs": n}, **res))

    if "verify" in only:
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

        def call():
            ok(lib.eth_kzg_verify_cell_kzg_proof_batch(H, C.c_uint64(N), pc, C.c_uint64(N), idx, C.c_uint64(N), pl, C.c_uint64(N), pp, C.byref(flag)))
            return bool(flag.value)
        call()
        t, good = best(call, args.reps)
        assert good is True
        b_cells[777 * CELL + CELL - 1] = bytes([cells_flat[777 * CELL + CELL - 1] ^ 1])
        tneg, okneg = best(call, 1)
        assert okneg is False
        commitments = [comms[:48]] * NCELLS
        t0 = time.perf_counter()
        okc = cref.verify_cell_kzg_proof_batch(commitments, list(range(NCELLS)), [cells_flat[k * CELL:(k + 1) * CELL] for k in range(NCELLS)],
                                               [proofs_flat[k * 48:(k + 1) * 48] for k in range(NCELLS)])
        tc = time.perf_counter() - t0
        assert okc is True
        emit({"config": "#5 verify_cell_kzg_proof_batch, 128 blobs x 128 cells in one call (the reference's pointer-array C ABI)", "metric": "cells/s",
              "value": N / t, "ms": 1e3 * t, "negative_case_ms": 1e3 * tneg, "cpu_port_cells_per_s_1thread": NCELLS / tc})
    if "pageable" in only:
        # config #3 through the batch entry point with PAGEABLE caller buffers (what a binding that does not pin its memory hands
        # over): the library stages through its own pinned memory, 64 blobs at a time, under the kernels
        n = 1024
        src = bytearray(syn.blobs(n))
        cells = bytearray(n * 262144)
        proofs = bytearray(n * 6144)
        status = bytearray(n)
        cb = (C.c_char * len(src)).from_buffer(src)
        cc = (C.c_char * len(cells)).from_buffer(cells)
        cp = (C.c_char * len(proofs)).from_buffer(proofs)
        cs = (C.c_char * len(status)).from_buffer(status)

        def call():
            ok(lib.eth_kzg_b200_compute_cells_and_kzg_proofs_batch(H, C.c_uint64(n), cb, cc, cp, cs))
            return True
        call()
        t, _ = best(call, args.reps)
        assert bytes(cells[:131072]) == bytes(src[:131072])
        emit({"config": "#3 compute_cells_and_kzg_proofs, batch of 1024 blobs, PAGEABLE host buffers through the C ABI", "metric": "blobs/s",
              "value": n / t, "ms": 1e3 * t})
    if "callers" in only:
        # the reference's own usage pattern: T host threads, each calling the SINGLE-blob ABI function in a loop on one shared
        # context (bindings/node/src/lib.rs:92-130 calls from the libuv pool).  The library coalesces concurrent callers.
        import threading
        T, per = 64, 4
        blobs = [syn.blob(9000 + i) for i in range(T)]

        def worker(i, barrier, res):
            cells = [C.create_string_buffer(2048) for _ in range(128)]
            proofs = [C.create_string_buffer(48) for _ in range(128)]
            pc = (C.c_void_p * 128)(*[C.addressof(b) for b in cells])
            pp = (C.c_void_p * 128)(*[C.addressof(b) for b in proofs])
            barrier.wait()
            for _ in range(per):
                r = lib.eth_kzg_compute_cells_and_kzg_proofs(H, blobs[i], pc, pp)
                if r.status != 0:
                    res[i] = "err"
                    return
            res[i] = proofs[127].raw

        def run_threads():
            barrier = threading.Barrier(T + 1)
            res = [None] * T
            th = [threading.Thread(target=worker, args=(i, barrier, res)) for i in range(T)]
            for t in th:
                t.start()
            barrier.wait()
            t0 = time.perf_counter()
            for t in th:
                t.join()
            return time.perf_counter() - t0, res
        run_threads()
        dt, res = run_threads()
        _, pf, _ = ctx.compute_cells_and_kzg_proofs_batch(b"".join(blobs), T)
        assert all(res[i] == pf[i * 6144 + 127 * 48:(i + 1) * 6144] for i in range(T)), "a coalesced caller got a wrong proof"
        emit({"config": "single-blob ABI calls (eth_kzg_compute_cells_and_kzg_proofs) from %d concurrent host threads on one context, coalesced by the library" % T,
              "metric": "blobs/s", "value": T * per / dt, "threads": T, "calls_per_thread": per, "ms_per_call_seen_by_a_thread": 1e3 * dt / per,
              "coalescing": "off" if os.environ.get("EKZG_NO_COALESCE") else "on"})
    ctx.close()


if __name__ == "__main__":
    main()
