#!/usr/bin/env python
"""bench.py -- round-trip (encode + decode) throughput of the bzip3 block codec on B200.

One "step" = bz3_encode_block + bz3_decode_block over every block of the workload (all blocks of a rank
in flight at once, one stream per block), i.e. one full round trip of the rank's data.

  value      MiB/s of uncompressed data through one round trip, inputs resident in HBM
             (bz3_b200_encode_resident_many / bz3_b200_decode_resident_many), device-timed.
  e2e        the same through the reference ABI bz3_encode_blocks / bz3_decode_blocks on pinned HOST
             buffers (H2D + D2H inside the timed region; for N>1 also the NCCL gather of the
             compressed blocks to rank 0).
  roofline   dominant kernel (by device time) vs the measured HBM peak, plus the suffix-sort radix passes.
  cpu_baseline / --impl reference : the unmodified reference (oracle/_ref/libbz3_ref.so, built from
             /root/reference by oracle/Makefile) on the host cores, same bytes, same block size.

Launch: python bench.py [--gpus N --steps K --warmup W]  or, for N>1,
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
"""
from __future__ import annotations

import os as _os

# One CUDA stream per block: give every stream its own hardware queue, otherwise a copy queued behind one
# block's seconds-long coder kernel falsely serialises the other blocks' short kernels (must be set before
# the CUDA context exists).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bzip3_b200 import synth  # noqa: E402

MIB = float(1 << 20)

WORKLOADS = {
    # BASELINE.json configs[1]: "enwik8-style 100 MB synthetic Zipf text, -b 16, 1 GPU"
    "zipf100m_b16": dict(gen="zipf_text", nbytes=100_000_000, block=16 << 20, seed=synth.SEED_ZIPF_TEXT,
                         desc="100 MB synthetic Zipf(1.1) text, -b 16 (6 blocks: 5 x 16 MiB + 16 113 920 B)"),
    # BASELINE.json configs[2]: "1 GiB synthetic source-code corpus, -b 256"
    "src1g_b256": dict(gen="source_corpus", nbytes=1 << 30, block=256 << 20, seed=synth.SEED_SOURCE,
                       desc="1 GiB synthetic source corpus, -b 256 (4 blocks)"),
    "src256m_b64": dict(gen="source_corpus", nbytes=256 << 20, block=64 << 20, seed=synth.SEED_SOURCE,
                        desc="256 MiB synthetic source corpus, -b 64 (4 blocks)"),
    # not a BASELINE config: many blocks per GPU, to measure what blocks in flight buy (the per-device
    # workspace pool lets a B200 hold far more blocks than it has SMs); 128 blocks = 128 single-CTA coder kernels at once
    "zipf2g_b16": dict(gen="zipf_text", nbytes=2 << 30, block=16 << 20, seed=synth.SEED_ZIPF_TEXT,
                       desc="2 GiB synthetic Zipf(1.1) text, -b 16 (128 blocks in flight per GPU)"),
    "zipf8m_b1": dict(gen="zipf_text", nbytes=8 << 20, block=1 << 20, seed=synth.SEED_ZIPF_TEXT,
                      desc="8 MiB synthetic Zipf text, -b 1 (8 blocks) -- quick self-test"),
}


def load_workload(name: str, rank: int):
    w = WORKLOADS[name]
    seed = w["seed"] + 1000 * rank
    cache_dir = os.environ.get("BZ3_B200_CACHE", "/tmp/bz3_b200_cache")
    path = os.path.join(cache_dir, f"{name}_{seed:x}.bin")
    data = None
    if os.path.exists(path) and os.path.getsize(path) == w["nbytes"]:
        data = np.fromfile(path, dtype=np.uint8)
    if data is None:
        data = getattr(synth, w["gen"])(w["nbytes"], seed=seed)
        try:
            os.makedirs(cache_dir, exist_ok=True)
            data.tofile(path)
        except OSError:
            pass
    bs = w["block"]
    blocks = [data[i:i + bs] for i in range(0, len(data), bs)]
    return blocks, bs, w


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- reference arm
def reference_lib():
    from tests import refs
    if refs.have_ref():
        return refs.ref(), "reference"
    return None, "port"


def reference_roundtrip(blocks, bs, jobs):
    """One round trip with the reference's own batch API (pthread per block, batches of `jobs` like src/main.c:352-378).
    Returns (seconds_encode, seconds_decode)."""
    from tests import refs
    R, kind = reference_lib()
    cap = refs.bound(bs) + 64
    if R is None:  # oracle port, single thread
        te = td = 0.0
        for b in blocks:
            buf = np.zeros(cap, np.uint8)
            buf[:len(b)] = b
            err = C.c_int8(0)
            t0 = time.perf_counter()
            r = refs.oracle().orc_encode_block(bs, refs.ptr(buf), len(b), C.byref(err))
            t1 = time.perf_counter()
            refs.oracle().orc_decode_block(bs, refs.ptr(buf), cap, r, len(b), C.byref(err))
            t2 = time.perf_counter()
            te += t1 - t0
            td += t2 - t1
            assert bytes(buf[:len(b)]) == b.tobytes()
        return te, td, kind
    states = [R.bz3_new(bs) for _ in range(min(jobs, len(blocks)))]
    bufs = [np.zeros(cap, np.uint8) for _ in states]
    te = td = 0.0
    try:
        for a in range(0, len(blocks), len(states)):
            grp = blocks[a:a + len(states)]
            n = len(grp)
            for b, buf in zip(grp, bufs):
                buf[:len(b)] = b
            hs = (C.c_void_p * n)(*states[:n])
            bp = (refs.u8p * n)(*[refs.ptr(x) for x in bufs[:n]])
            sz = (C.c_int32 * n)(*[len(b) for b in grp])
            t0 = time.perf_counter()
            R.bz3_encode_blocks(hs, bp, sz, n)
            t1 = time.perf_counter()
            bsz = (C.c_size_t * n)(*[cap] * n)
            osz = (C.c_int32 * n)(*[len(b) for b in grp])
            R.bz3_decode_blocks(hs, bp, bsz, sz, osz, n)
            t2 = time.perf_counter()
            te += t1 - t0
            td += t2 - t1
            for b, buf, st in zip(grp, bufs, states):
                assert R.bz3_last_error(st) == 0 and bytes(buf[:len(b)]) == b.tobytes()
    finally:
        for st in states:
            R.bz3_free(st)
    return te, td, kind


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    blocks, bs, w = load_workload(args.workload, 0)
    total = sum(len(b) for b in blocks)
    jobs = min(host_threads(), 64)
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        te, td, kind = reference_roundtrip(blocks, bs, jobs)
        if i >= args.warmup:
            times.append((te, td))
    te = sum(t[0] for t in times) / len(times)
    td = sum(t[1] for t in times) / len(times)
    ms = (te + td) * 1e3
    val = total / MIB / (te + td)
    out = {
        "impl": "reference", "metric": "roundtrip_MiB_per_s", "value": round(val, 3), "unit": "MiB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "description": w["desc"], "block_size": bs, "blocks": len(blocks),
                   "bytes": total},
        "encode_MiB_per_s": round(total / MIB / te, 3), "decode_MiB_per_s": round(total / MIB / td, 3),
        "cpu_baseline": {"value": round(val, 3), "unit": "MiB/s", "cores": min(jobs, len(blocks)), "kind": kind,
                         "sample": f"full workload, {len(blocks)} blocks, bz3_encode_blocks/bz3_decode_blocks with "
                                   f"{min(jobs, len(blocks))} pthreads (host has {host_threads()} usable cores)"},
        "e2e": {"value": round(val, 3), "unit": "MiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(out)


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200_arm(args, rank, world, local_rank):
    import torch
    import bzip3_b200
    from tests import refs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the codec has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L = bzip3_b200.lib()
    blocks, bs, w = load_workload(args.workload, rank)
    nb = len(blocks)
    total = sum(len(b) for b in blocks)
    states = [bzip3_b200.Bz3State(bs) for _ in blocks]
    hs = (C.c_void_p * nb)(*[s.handle for s in states])
    cm_variants = (L.bz3_b200_get_variant(states[0].handle, 5 + 100), L.bz3_b200_get_variant(states[0].handle, 5 + 200))
    lzp_variant = L.bz3_b200_get_variant(states[0].handle, 3)
    sizes = (C.c_int32 * nb)(*[len(b) for b in blocks])
    osz = (C.c_int32 * nb)(*[len(b) for b in blocks])
    enc_sizes = (C.c_int32 * nb)()
    dec_res = (C.c_int32 * nb)()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- stage inputs in HBM (untimed) and check the round trip once against the oracle/reference
    for s, b in zip(states, blocks):
        assert L.bz3_b200_upload(s.handle, refs.ptr(b), len(b)) == 0
    L.bz3_b200_encode_resident_many(hs, sizes, enc_sizes, nb)
    assert all(e > 0 for e in enc_sizes), list(enc_sizes)
    if rank == 0:  # bit-exactness gate on block 0 (full check lives in tests/)
        got = np.zeros(enc_sizes[0], np.uint8)
        assert L.bz3_b200_download(states[0].handle, refs.ptr(got), enc_sizes[0]) == 0
        R, kind = reference_lib()
        if R is not None and len(blocks[0]) <= (64 << 20):
            want = refs.api_encode_block(R, blocks[0].tobytes(), bs)[0]
            assert want == got.tobytes(), "block 0 differs from the reference encoder"
    L.bz3_b200_decode_resident_many(hs, enc_sizes, osz, dec_res, nb)
    assert list(dec_res) == [len(b) for b in blocks], list(dec_res)
    chk = np.zeros(len(blocks[-1]), np.uint8)
    assert L.bz3_b200_download(states[-1].handle, refs.ptr(chk), len(chk)) == 0
    assert chk.tobytes() == blocks[-1].tobytes(), "round trip mismatch"

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def resident_step():
        flush.fill_(1)
        barrier()
        ev[0].record()
        L.bz3_b200_encode_resident_many(hs, sizes, enc_sizes, nb)
        ev[1].record()
        L.bz3_b200_decode_resident_many(hs, enc_sizes, osz, dec_res, nb)
        ev[2].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])

    for _ in range(args.warmup):
        resident_step()
    for s in states:
        s.stats_reset()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    enc_ms, dec_ms = [], []
    for _ in range(args.steps):
        e, d = resident_step()
        enc_ms.append(e)
        dec_ms.append(d)
    # read the kernels in effect again: a promoted kernel may have been retired during the warm-up
    cm_variants = (L.bz3_b200_get_variant(states[0].handle, 5 + 100), L.bz3_b200_get_variant(states[0].handle, 5 + 200))
    lzp_variant = L.bz3_b200_get_variant(states[0].handle, 3)
    launches = sum(s.launches() for s in states)
    stage_enc = [s.stage_ms(False) for s in states]
    stage_dec = [s.stage_ms(True) for s in states]
    sort_records = 0
    for s in states:
        rec = C.c_uint64(0)
        L.bz3_b200_last_sort_stats(s.handle, C.byref(rec), None, None)
        sort_records += rec.value
    step_ms = [a + b for a, b in zip(enc_ms, dec_ms)]
    t = torch.tensor([sum(step_ms), sum(enc_ms), sum(dec_ms)], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(total)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    sum_ms, sum_enc, sum_dec = [float(x) for x in t.tolist()]
    job_bytes = float(tot.item())
    ms_per_step = sum_ms / args.steps
    value = job_bytes / MIB / (ms_per_step / 1e3)

    # ---------------- end to end through the reference ABI with pinned host buffers
    cap = bzip3_b200.bound(bs) + 64
    pinned = [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in blocks]
    for p, b in zip(pinned, blocks):
        p[:len(b)] = torch.from_numpy(b)
    bp = (refs.u8p * nb)(*[C.cast(p.data_ptr(), refs.u8p) for p in pinned])
    bsz = (C.c_size_t * nb)(*[cap] * nb)
    e2e_ms = []
    h2d = d2h = 0
    # ---------------- isolated suffix-sort leg: one block alone on the device, radix passes bracketed by CUDA events
    sa_leg = None
    if rank == 0:
        st0 = states[0]
        tmp_out = np.zeros(len(blocks[0]) + 64, np.uint8)
        for rep in range(4):
            if rep == 1:
                st0.stats_reset()
            flush.fill_(1)
            torch.cuda.synchronize()
            L.bz3_b200_stage_bwt(st0.handle, refs.ptr(blocks[0]), len(blocks[0]), refs.ptr(tmp_out))
        rec, rounds, sms = C.c_uint64(0), C.c_int32(0), C.c_double(0)
        L.bz3_b200_last_sort_stats(st0.handle, C.byref(rec), C.byref(rounds), C.byref(sms))
        sa_leg = {"records_x_passes": rec.value / 3, "sort_ms": sms.value / 3, "rounds": rounds.value,
                  "block_bytes": len(blocks[0])}
    e2e_warm = 1  # the device is already warm from the resident leg
    for i in range(e2e_warm + args.steps):
        flush.fill_(1)
        barrier()
        ev[0].record()
        csz = (C.c_int32 * nb)(*[len(b) for b in blocks])
        L.bz3_encode_blocks(hs, bp, csz, nb)
        comp = [int(c) for c in csz]
        if dist:  # ordered gather of the variable-length compressed blocks to rank 0 over NCCL
            szt = torch.tensor(comp, dtype=torch.int32, device="cuda")
            allsz = [torch.empty_like(szt) for _ in range(world)]
            dist.all_gather(allsz, szt)
            mx = int(max(int(a.max()) for a in allsz))
            payload = torch.zeros((nb, mx), dtype=torch.uint8, device="cuda")
            for k, (p, c) in enumerate(zip(pinned, comp)):
                payload[k, :c].copy_(p[:c], non_blocking=True)
            gathered = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
            dist.gather(payload, gathered, dst=0)
        L.bz3_decode_blocks(hs, bp, bsz, csz, osz, nb)
        ev[1].record()
        torch.cuda.synchronize()
        assert all(s.last_error == 0 for s in states)
        if i >= e2e_warm:
            e2e_ms.append(ev[0].elapsed_time(ev[1]))
            h2d = total + sum(comp)
            d2h = sum(comp) + total
    assert bytes(pinned[0][:len(blocks[0])].numpy()) == blocks[0].tobytes()
    t2 = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_step_ms = float(t2.item()) / args.steps
    e2e_val = job_bytes / MIB / (e2e_step_ms / 1e3)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

        def stage_total(st, name):
            return sum(d[name] for d in st)

        # dominant kernel = the CM coder of the slower direction; one launch per block per step
        cm_dec_ms = stage_total(stage_dec, "cm") / (nb * args.steps)
        cm_enc_ms = stage_total(stage_enc, "cm") / (nb * args.steps)
        avg_n = total / nb
        avg_c = sum(int(e) for e in enc_sizes) / nb
        dec_names = {0: "cm_decode_tree_kernel", 1: "cm_decode_single_kernel", 3: "cm_decode_paths_kernel",
                     4: "cm_decode_lanes_kernel", 5: "cm_decode_paths2_kernel", 6: "cm_decode_walkers_kernel"}
        enc_names = {1: "cm_encode_single_kernel", 2: "cm_encode_chunked_kernel<1>", 4: "cm_encode_chunked_kernel<2>",
                     6: "cm_encode_chunked_kernel<3>"}
        v_enc, v_dec = cm_variants
        dom = (dec_names.get(v_dec, "cm_decode_tree_kernel") if cm_dec_ms >= cm_enc_ms
               else enc_names.get(v_enc, "cm_encode_chunked_kernel<0>"))
        dom_ms = max(cm_dec_ms, cm_enc_ms)
        dom_bytes = avg_n + avg_c  # the coder reads/writes the BWT bytes once and the payload once
        achieved = dom_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        sa_bytes = 32.0 * sa_leg["records_x_passes"]  # 8 (hist) + 12 + 12 bytes per record per radix pass
        sa_ach = sa_bytes / (sa_leg["sort_ms"] / 1e3) / 1e9 if sa_leg["sort_ms"] > 0 else 0.0
        whole = (16.0 * total + sum(int(e) for e in enc_sizes)) * 2 / (ms_per_step / 1e3) / 1e9
        cpu = None
        try:
            jobs = min(host_threads(), 64)
            te, td, kind = reference_roundtrip(blocks, bs, jobs)
            cpu = {"value": round(total / MIB / (te + td), 3), "unit": "MiB/s", "cores": min(jobs, nb), "kind": kind,
                   "sample": f"one round trip of the full rank-0 workload ({nb} blocks, {total} B) with the reference's "
                             f"bz3_encode_blocks/bz3_decode_blocks, {min(jobs, nb)} pthreads of {host_threads()} usable cores",
                   "encode_MiB_per_s": round(total / MIB / te, 3), "decode_MiB_per_s": round(total / MIB / td, 3)}
        except Exception as ex:  # the CPU leg must never take the GPU numbers down with it
            cpu = {"value": None, "unit": "MiB/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        out = {
            "metric": "roundtrip_MiB_per_s", "value": round(value, 3), "unit": "MiB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "block_size": bs, "blocks_per_gpu": nb,
                       "bytes_per_gpu": total, "parallelism": f"blocks sharded {world} way(s), one stream per block",
                       "l2": "256 MiB flush buffer written before every timed step; block buffers + shared stage workspaces >> 126 MB L2",
                       "kernels_in_effect": {"entropy_encoder": int(cm_variants[0]), "entropy_decoder": int(cm_variants[1]),
                                             "lzp": int(lzp_variant), "retired_after_checksum_failure": int(L.bz3_b200_demotions()),
                                             "how": "on-device self-test at the first bz3_new (DESIGN.md 6c); 0/0/3 = round-1 kernels"},
                       "hbm_bytes": {"per_block_state": int(L.bz3_b200_device_bytes(states[0].handle)),
                                     "shared_stage_workspaces": int(L.bz3_b200_workspace_bytes(states[0].handle))},
                       "definition": "one step = encode + decode of every block; value = bytes / step time"},
            "encode_MiB_per_s": round(job_bytes / MIB / (sum_enc / args.steps / 1e3), 3),
            "decode_MiB_per_s": round(job_bytes / MIB / (sum_dec / args.steps / 1e3), 3),
            "compressed_bytes_rank0": int(sum(int(e) for e in enc_sizes)),
            "e2e": {"value": round(e2e_val, 3), "unit": "MiB/s", "ms_per_step": round(e2e_step_ms, 3),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "bz3_encode_blocks + bz3_decode_blocks on pinned host buffers"
                           + (" + NCCL gather of compressed blocks to rank 0" if world > 1 else "")},
            "gpu_launches": int(launches),
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 6), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 9), "traffic": None, "peak_source": peak_src,
                         "note": "serial range-coder recurrence: latency-bound, not bandwidth-bound (DESIGN.md)",
                         "avg_launch_ms": round(dom_ms, 3), "algorithmic_bytes_per_launch": int(dom_bytes)},
            "roofline_sa_radix": {"kernel": "rs_tile_hist_kernel+rs_scatter_kernel (suffix-sort radix passes)",
                                  "bound": "hbm", "achieved": round(sa_ach, 3), "peak": hbm_peak, "unit": "GB/s",
                                  "frac": round(sa_ach / hbm_peak, 6),
                                  "algorithmic_bytes": int(sa_bytes), "ms": round(sa_leg["sort_ms"], 3),
                                  "rounds": sa_leg["rounds"], "block_bytes": sa_leg["block_bytes"],
                                  "note": "one block alone on the device; CUDA events around every radix sort of the "
                                          "suffix sorter (tile histogram + scan + scatter launches, no host sync inside); "
                                          "32 B per record per 8-bit pass = 8 (hist read) + 12 (read) + 12 (write)"},
            "roofline_whole_job": {"achieved": round(whole, 3), "peak": hbm_peak, "unit": "GB/s",
                                   "frac": round(whole / hbm_peak, 9), "note": "SURVEY 8(d): (16 n + c) bytes per direction"},
            "stage_ms_per_step": {"encode": {k: round(stage_total(stage_enc, k) / args.steps, 3) for k in bzip3_b200.STAGES},
                                  "decode": {k: round(stage_total(stage_dec, k) / args.steps, 3) for k in bzip3_b200.STAGES},
                                  "note": "summed over the rank's concurrently running blocks"},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        _emit(out)
    for s in states:
        s.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the duration of the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(obj):
    line = json.dumps(obj)
    if _REAL_STDOUT is not None:
        _REAL_STDOUT.write(line + "\n")
        _REAL_STDOUT.flush()
    else:
        print(line, flush=True)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BZ3_BENCH_WORKLOAD", "zipf100m_b16"), choices=sorted(WORKLOADS))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
