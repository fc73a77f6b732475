#!/usr/bin/env python
"""slab_bench.py -- one large grid slab-decomposed over the GPUs of a node (one process per GPU, NCCL).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/slab_bench.py --grid 16384 --steps 10 --warmup 2 [--check 1024] [--chunks 4]

--check G : first verify the slab path against the single-GPU path on a G x G Kuo-2004 field
            (rank 0 gathers the rows; relative L2 must be < 2e-6).
Prints ONE JSON line on rank 0: grid-pt*steps/s of the whole job (max over ranks, CUDA events), the
all-to-all time measured on the communication stream and the NVLink roofline of SURVEY.md 8(e).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--field", default="const", choices=["const", "elliptic"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    res = run(args)
    if res is not None:
        print(json.dumps(res))


def run(args):
    """-> result dict on rank 0, None elsewhere (also used by `bench.py --slab`)"""

    import torch
    import torch.distributed as dist
    import fields
    import xlab_fftbarotropic_b200 as xfb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def new_id():
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(xfb.nccl_unique_id()), dtype=torch.uint8))
        if world > 1:
            dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    result = {"n_gpus": world, "chunks": args.chunks}

    # ---- correctness against the single-GPU path ---------------------------------------------------
    if args.check:
        n = args.check
        v0 = fields.kuo2004(n)
        sb = xfb.SlabBackend(n, rank, world, new_id(), nchunks=args.chunks, device=local_rank)
        r0 = rank * (n // world)
        sb.set_vorticity(v0[r0:r0 + n // world])
        sb.step(3, 3.0)
        errs = {}
        for name, which in (("vort", xfb.capi.VORT), ("u", xfb.capi.U), ("psi", xfb.capi.PSI)):
            mine = torch.from_numpy(sb.get_field(which)).to(dev)
            if world > 1:
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)
                full = torch.cat(parts, 0).cpu().numpy()
            else:
                full = mine.cpu().numpy()
            if rank == 0:
                one = xfb.Backend(n, device=local_rank)
                one.set_vorticity(v0)
                one.step(3, 3.0)
                ref = one.get_field(which)
                one.close()
                errs[name] = float(np.linalg.norm((full - ref).astype(np.float64)) / np.linalg.norm(ref.astype(np.float64)))
        sb.close()
        barrier()
        if rank == 0:
            result["check"] = {"grid": n, "rel_l2_vs_single_gpu": errs, "ok": all(e < 2e-6 for e in errs.values())}

    # ---- throughput ----------------------------------------------------------------------------------
    n = args.grid
    gen = fields.const_vortex if args.field == "const" else fields.elliptic
    dt = 3.0 if args.field == "const" else float(np.float32(0.6 * 2.83 / (0.943 * np.pi * n / 600000.0 * 83.0)))
    rows = n // world
    r0 = rank * rows
    if rank == 0:
        v0 = gen(n)
        chunks = [torch.from_numpy(np.ascontiguousarray(v0[q * rows:(q + 1) * rows])) for q in range(world)]
    mine = torch.empty((rows, n), dtype=torch.float32, device=dev)
    if world > 1:
        for q in range(world):                      # rank 0 generated the field; hand every rank its rows
            if q == 0:
                if rank == 0:
                    mine.copy_(chunks[0])
            elif rank == 0:
                dist.send(chunks[q].to(dev), q)
            elif rank == q:
                dist.recv(mine, 0)
    else:
        mine.copy_(chunks[0])
    sb = xfb.SlabBackend(n, rank, world, new_id(), nchunks=args.chunks, device=local_rank)
    transport = sb.transport
    sb.set_vorticity(int(mine.data_ptr()))
    sb.step(args.warmup, dt)
    sb.sync()
    stream = torch.cuda.ExternalStream(sb.stream, device=local_rank)
    sb.profile(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sb.launch_count
    import time
    with torch.cuda.stream(stream):
        e0.record(stream)
        h0 = time.perf_counter()
        sb.step(args.steps, dt)
        host_enqueue_ms = (time.perf_counter() - h0) * 1e3
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    prof = sb.profile_read()
    a2a = sb.a2a_read() if world > 1 else {"a2a_ms": 0.0, "exchanges": 0}
    launches = sb.launch_count - l0
    if world > 1:
        t = torch.tensor([ms, a2a["a2a_ms"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, a2a_ms = float(t[0]), float(t[1])
    else:
        a2a_ms = 0.0
    out = sb.get_field(xfb.capi.VORT)
    finite = bool(np.isfinite(out).all())
    # end to end through the C ABI with HOST buffers: every step uploads this rank's rows and reads them back
    e2e_steps = max(1, getattr(args, "e2e_steps", 3))
    host_in = mine.cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    hin, hout = int(host_in.data_ptr()), int(host_out.data_ptr())
    sb.set_vorticity(host_in.numpy())
    sb.step(1, dt)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(e2e_steps):
            sb._ck(sb._L.xfb_set_vorticity(sb._h, 0, hin))
            sb.step(1, dt)
            sb._ck(sb._L.xfb_get_field(sb._h, 0, xfb.capi.VORT, hout))
        e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t2 = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_e2e = float(t2[0])
    sb.close()
    if rank == 0:
        G = float(n) * n
        value = G * args.steps / (ms * 1e-3)
        # bytes sent per GPU per step = 20 * 4 N^2 / P * (P-1)/P  (spectral array ~ 4 B per grid point)
        nv_bytes = 20.0 * 4.0 * G / world * (world - 1) / world if world > 1 else 0.0
        result.update({
            "metric": "rk4_grid_point_steps_per_s", "value": value, "unit": "grid-pt*steps/s", "grid": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "scaling": "strong", "field": args.field, "dt": dt,
            "state_finite": finite, "gpu_launches_per_rank": int(launches), "transport": transport, "host_enqueue_ms_per_step": host_enqueue_ms / args.steps,
            "hbm_roofline_frac": 240.0 * value / world / 1e9 / 6455.9,
            "a2a": {"ms_per_step_on_comm_stream": a2a_ms / args.steps, "exchanges": a2a["exchanges"],
                    "nvlink_bytes_per_gpu_per_step": nv_bytes,
                    "achieved_gbs_per_direction": (nv_bytes / (a2a_ms / args.steps * 1e-3) / 1e9) if a2a_ms > 0 else None,
                    "peak_gbs": 770.0, "peak_source": "B200_PROFILING.md peer copy"},
            "kernels": {"row_ms_per_step": prof["row_ms"] / args.steps, "col_ms_per_step": prof["col_ms"] / args.steps,
                        "note": "row/col spans include the wait for their exchanges"},
            "e2e": {"value": G * e2e_steps / (ms_e2e * 1e-3), "unit": "grid-pt*steps/s", "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps, "h2d_bytes_per_step": 4 * G, "d2h_bytes_per_step": 4 * G},
        })
    if world > 1:
        dist.destroy_process_group()
    return result if rank == 0 else None


if __name__ == "__main__":
    main()
