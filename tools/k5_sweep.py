#!/usr/bin/env python3
"""K5 (the two G1 NTTs) stage time against batch size: radix-2 kernel, the radix-4 latency-mode kernel (EKZG_K5_R4_MAX) and the
latter with cooperative multiplication units, four lanes per field element (EKZG_K5_COOP_MAX): where the switch-overs belong.  One JSON line per (batch size, form)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__  # noqa: E402

pkg = __graft_entry__.load_package()
import importlib  # noqa: E402
syn = importlib.import_module("eth_kzg_b200.synthetic")
ctx = pkg.DASContext(use_precomp=True)
names = ["K1", "K2", "K4", "K5", "K6"]
stream = torch.cuda.current_stream()
blobs_all = syn.blobs(512)
for n in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,8,16,32,48,64,96,128,160,192,224,256,512".split(","))]:
    d_in = torch.frombuffer(bytearray(blobs_all[:n * 131072]), dtype=torch.uint8).cuda()
    d_cells = torch.empty(n * 262144, dtype=torch.uint8, device="cuda")
    d_proofs = torch.empty(n * 6144, dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(n, dtype=torch.int32, device="cuda")
    ref = None
    os.environ["EKZG_DIRECT_MAX"] = "0"     # (one or two blobs would otherwise bypass the G1 transforms)
    for form, r4max, coopmax in (("radix2", "0", "0"), ("radix4", "256", "0"), ("radix4_coop", "256", "256")):
        if form != "radix2" and n > 256:
            continue
        os.environ["EKZG_K5_R4_MAX"] = r4max
        os.environ["EKZG_K5_COOP_MAX"] = coopmax
        step = lambda: ctx.compute_cells_and_kzg_proofs_device(n, d_in.data_ptr(), d_cells.data_ptr(), d_proofs.data_ptr(), d_status.data_ptr(), stream.cuda_stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ctx.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        nb, ms = ctx.collect_stage_times()
        ctx.set_profiling(False)
        out = d_proofs.clone()
        if ref is None:
            ref = out
        same = bool(torch.equal(ref, out))
        print(json.dumps({"blobs": n, "k5": form, "step_ms": e0.elapsed_time(e1) / 5, "stages_ms": {k: round(v / max(nb, 1), 3) for k, v in zip(names, ms)}, "same_as_radix2": same}), flush=True)
ctx.close()
