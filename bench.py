#!/usr/bin/env python
"""bench.py -- matched frames/s of MANet's matching + map-memory hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = what the propagation loop does per frame before the segmentation head
(test.py:237-259 -> IntVOS.prop_seghead, IntVOS.py:600-661): global matching of the current
frame against the annotated frame (+ normalisation + global-map memory min), local matching
against the previous frame (+ local-map memory store/select).  Workload: synthetic DAVIS-480p
embeddings (C=100, 120x214 at stride 4), 5 objects (N=6 ids), max_distance=12 (the reference
default, config.py:50), k=1, every reference pixel labelled (R = M = 25 680; cfg.TEST_MODE off: no
unlabelled pixels to drop -- the scribble regime is what the propagation_50 / session_8_rounds legs run).

  value : frames/s with inputs resident in HBM (CUDA events on the launching stream, L2 flushed
          between steps), whole job over all ranks (independent sequences per GPU: weak scaling)
  e2e   : the same step through the C-ABI session with HOST buffers (pinned H2D of the three
          embeddings + two label maps, D2H of both maps, inside the timed region)
  --impl reference : the reference's CPU torch path (oracle port, all host threads), bounded sample

Nothing here reads /root/reference.  oracle/ is used only for the CPU baseline legs.
"""
from __future__ import annotations

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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, C, N_IDS, D_LOCAL = 120, 214, 100, 6, 12
M_PIX = H * W
WORKLOAD = "global+local matching + map-memory update, 480p emb 100x120x214, 5 objects (N=6), max_distance=12, k=1"
ALGO_FLOP_GLOBAL = 2.0 * M_PIX * M_PIX * C                      # 2*M*R*C, SURVEY.md section 8d
ALGO_BYTES_LOCAL = 4.0 * (2 * C * H * W + H * W + H * W * N_IDS)
# executed tensor-core work: M padded to 256, R padded per 256-row bucket, K steps of 16.  Filter-and-refine engine: the ONE
# product qh.rh = 7 K steps per tile (6 x 16 channels + the folded step with the last 4 channels and the bias); the
# three-product engine (MANET_GM_ENGINE=exact3) runs 19 (7 + 6 + 6).
GM_EXACT3 = os.environ.get("MANET_GM_ENGINE", "")[:1] in ("3", "e")
EXEC_FLOP_GLOBAL = 2.0 * (101 * 256) * (103 * 256) * 16 * (19 if GM_EXACT3 else 7)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops"], "tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json, bf16 burst)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = sorted(sm)[len(sm) // 2:] if sm else []     # upper half ~ samples under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_inputs(seed):
    import torch
    gen = torch.Generator().manual_seed(seed)
    mk = lambda: 0.1 * torch.relu(torch.randn(C, H, W, generator=gen))
    ref, prev = mk(), mk()
    cur = prev + 0.02 * torch.randn(C, H, W, generator=gen)
    blob = lambda: torch.randint(0, N_IDS, (H // 8, W // 8 + 1), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int().contiguous()
    return ref, prev, cur, blob(), blob()


# ------------------------------------------------------------------------------------------ CPU legs
CPU_KIND_NOTE = ("port: oracle/manet_oracle.py, the op-for-op torch-CPU restatement of IntVOS.py:23-434,600-661 (bit-exact against "
                 "the unmodified reference on the committed goldens); the reference's own Python sources live under /root/reference, "
                 "which does not exist on the GPU box, so oracle/ref_shim.py cannot run here")


def cpu_reference_step(inputs):
    """One WHOLE frame of the reference CPU path (oracle port), nothing extrapolated: global matching with the reference's
    own 10 query chunks (IntVOS.py:139-152, n_chunks=10 at :610) + normalisation + global-map memory min, local matching
    (d=12) + local-map memory store/select.  Returns (seconds, global seconds, local seconds)."""
    import torch
    from oracle import manet_oracle as O
    ref, prev, cur, ref_lab, prev_lab = inputs
    gmem, lmem = {}, ({}, {})
    t0 = time.perf_counter()
    refv, curv, prevv = ref.permute(1, 2, 0), cur.permute(1, 2, 0), prev.permute(1, 2, 0)
    g, ids = O.global_match(refv, curv, ref_lab.unsqueeze(-1), 1, torch.tensor(N_IDS - 1), n_chunks=10, test_mode=True)
    g = O.global_map_read_update(gmem, "s", 1, O.normalize_distance(g))
    t1 = time.perf_counter()
    loc = O.local_match(prevv, curv, prev_lab.unsqueeze(-1), ids, D_LOCAL)
    loc, _ = O.local_map_store_select(lmem, "s", 1, 1, 0, loc)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = synth_inputs(0)
    for _ in range(args.warmup):
        cpu_reference_step(inputs)
    times, tg, tl = [], [], []
    for _ in range(args.steps):
        t, a, b = cpu_reference_step(inputs)
        times.append(t); tg.append(a); tl.append(b)
    per_frame = sum(times) / len(times)
    value = 1.0 / per_frame
    desc = ("every step is one whole 480p frame (25 680 queries x 25 680 references in the reference's 10 chunks, N=6; local d=12; "
            "both memory updates), no extrapolation; " + CPU_KIND_NOTE)
    line = {"impl": "reference", "metric": "matched frames/sec (global+local, 480p, 5 obj)", "value": value,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_frame * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc,
                             "global_s_per_frame": sum(tg) / len(tg), "local_s_per_frame": sum(tl) / len(tl)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """Bounded CPU sample for the own-arm line (rank 0, N=1): one warm-up frame + 3 timed whole frames."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = synth_inputs(0)
    cpu_reference_step(inputs)
    runs = [cpu_reference_step(inputs) for _ in range(3)]
    t = sum(r[0] for r in runs) / 3
    return {"value": 1.0 / t, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "1 warm-up + 3 timed WHOLE frames of the same workload (mean), no extrapolation; " + CPU_KIND_NOTE,
            "global_s_per_frame": sum(r[1] for r in runs) / 3, "local_s_per_frame": sum(r[2] for r in runs) / 3}


# ------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs local to its GPU BEFORE the session allocates its pinned staging buffers (first touch places
    them).  With N ranks each streaming ~10 MB per step through host memory, staging buffers on the far NUMA node were the
    e2e limiter at N = 8.  Multi-rank runs only: the single-rank CPU baseline leg wants every core."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"pci": bus, "numa_node": node, "cpus_bound": len(allowed)}
    except Exception as e:      # no sysfs / attribute: run unbound
        return {"error": str(e)[:120]}


def run_own_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from cvpr2020_manet_b200 import _lib, engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    _lib.check(L.manet_check_device(), "device check")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    sess = engine.MatchingSession(H, W, C, N_IDS, D_LOCAL, n_frames=104)
    ref, prev, cur, ref_lab, prev_lab = synth_inputs(1000 + rank)
    sess.ref[:], sess.prev[:], sess.cur[:] = ref.numpy(), prev.numpy(), cur.numpy()
    sess.ref_labels[:], sess.prev_labels[:] = ref_lab.numpy(), prev_lab.numpy()
    sess.upload(); sess.sync()
    stream = torch.cuda.ExternalStream(sess.stream, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    K, Wm = args.steps, args.warmup

    def timed_pass(serial, ref_cache=True):
        """K steps, one CUDA-event pair per step on the launching stream, L2 flushed between steps."""
        for i in range(Wm):
            sess.step_device(1 + i % 100, 1, 0, drop_unlabelled=False, serial=serial, ref_cache=ref_cache)
        sess.sync()
        L.manet_profile_reset()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        barrier()
        w0 = time.perf_counter()
        with torch.cuda.stream(stream):
            for i in range(K):
                flush_buf.fill_(i & 0xFF)                      # L2 flush, outside the timed events
                starts[i].record(stream)
                sess.step_device(1 + (Wm + i) % 100, 1, 0, drop_unlabelled=False, serial=serial, ref_cache=ref_cache)
                stops[i].record(stream)
        sess.sync()
        barrier()
        wall = time.perf_counter() - w0
        return sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / 1e3, wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # pass 1 (headline): local branch forked onto a second stream.  pass 2: everything on one stream, with the
    # library's per-kernel events switched on (they sit between the kernels, so they stay out of the headline pass;
    # on one stream their timings are not inflated by the two branches queueing for the same SMs).
    L.manet_profile_reset_launches()
    total_s, wall_dev = timed_pass(serial=False)
    launches_total = int(L.manet_profile_launch_count())          # warm-up + timed steps of the headline pass, counted by the library
    gpu_launches = launches_total * K // (K + Wm)                  # every step launches the same kernels
    nocache_s, _ = timed_pass(serial=False, ref_cache=False)      # every step rebuilds the reference side (a sequence's first frame)
    L.manet_profile_enable(K + 4)
    serial_s, _ = timed_pass(serial=True)
    # kernel timings recorded inside the library on the launching stream (serial pass)
    prof = {}
    for slot, name in ((0, "global_tcgen05"), (1, "local_main"), (2, "local_prepass"), (3, "global_refine"), (4, "global_rescan"),
                       (5, "global_prepass"), (6, "global_exact3")):
        buf = (ctypes.c_float * (K + 4))()
        n = ctypes.c_int(0)
        _lib.check(L.manet_profile_read(slot, buf, K + 4, ctypes.byref(n)), "manet_profile_read")
        vals = [buf[i] for i in range(n.value)]
        prof[name] = (sum(vals) / len(vals)) if vals else None
    L.manet_profile_enable(0)

    # end-to-end through the host-buffer C-ABI session.  Every step copies its three embeddings and two
    # label maps host->device from pinned memory and both result maps device->host; the session's two
    # slots let the upload of step i+1 overlap the kernels of step i (manet_session_submit_host/_wait).
    for k in ("ref", "prev", "cur", "ref_labels", "prev_labels"):
        sess.slots[1][k][:] = sess.slots[0][k]
    for _ in range(min(3, Wm)):
        sess.step_host(50, 1, 0, drop_unlabelled=False)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        sess.step_host(1 + i % 100, 1, 0, drop_unlabelled=False)
    e2e_sync_s = time.perf_counter() - t0
    barrier()
    checksum = 0.0
    t0 = time.perf_counter()
    for i in range(min(2, K)):
        sess.submit_host(i % 2, 1 + i % 100, 1, 0, drop_unlabelled=False)
    for i in range(K):
        og, ol = sess.wait(i % 2)
        checksum += float(og[0, 0, 0]) + float(ol[-1, -1, -1])      # the host reads the step's result
        if i + 2 < K:
            sess.submit_host(i % 2, 1 + (i + 2) % 100, 1, 0, drop_unlabelled=False)
    e2e_s = time.perf_counter() - t0
    barrier()
    # streaming propagation (MANET_STEP_STREAM) -- the headline e2e: the sequence starts in the untimed warm-up (its first
    # step uploads the annotated frame, its scribble labels and the first previous frame once, as a propagation does at
    # test.py:237); every timed step uploads its own new inputs (the new frame's embedding + the previous frame's labels)
    # from pinned host memory and downloads both result maps.
    for i in range(Wm):
        sess.submit_host(i % 2, 1 + i % 100, 1, 0, drop_unlabelled=False, stream=True, reset=(i == 0))
        sess.wait(i % 2)
    barrier()
    t0 = time.perf_counter()
    for i in range(min(2, K)):
        sess.submit_host((Wm + i) % 2, 1 + (Wm + i) % 100, 1, 0, drop_unlabelled=False, stream=True)
    for i in range(K):
        og, ol = sess.wait((Wm + i) % 2)
        checksum += float(og[0, 0, 0]) + float(ol[-1, -1, -1])
        if i + 2 < K:
            sess.submit_host((Wm + i) % 2, 1 + (Wm + i + 2) % 100, 1, 0, drop_unlabelled=False, stream=True)
    e2e_stream_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([total_s, e2e_s, e2e_sync_s, serial_s, e2e_stream_s, nocache_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_s, e2e_s, e2e_sync_s, serial_s, e2e_stream_s, nocache_s = (float(t[i]) for i in range(6))

    # the head / propagation legs describe one GPU; multi-GPU runs (independent sequences) report the scaling metric only.
    # They run BEFORE the 1080p sharded-matching leg: right after that leg's ~0.2 s of back-to-back 28 ms tensor-core kernels the
    # head's own leg read 0.81 ms instead of 0.62 ms (power management still settling), while the propagation leg a second later
    # was unaffected.
    extra = world == 1 and os.environ.get("MANET_BENCH_SEGHEAD", "1") == "1"
    seghead = seghead_leg(dev) if extra else None
    propagation = propagation_leg(dev) if extra else None
    session = session_leg(dev) if extra else None
    intvos_fwd = intvos_forward_leg(dev) if extra else None
    sharded = sharded_1080p_leg(dev, rank, world) if os.environ.get("MANET_BENCH_SHARDED", "1") == "1" else None

    if rank == 0:
        peaks = load_peaks()
        k_ms = prof["global_tcgen05"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        tfile = json.load(open(tpath)) if os.path.exists(tpath) else {}
        traffic = tfile.get("global_exact3_dram_bytes_per_launch" if GM_EXACT3 else "global_tcgen05_dram_bytes_per_launch")
        lm_inst = tfile.get("lm_umma_inst_executed")
        # slot 6 = the three-product chain: enqueued as well, exits at once when the pre-pass picked filter-and-refine (and vice versa)
        r_ms = (prof.get("global_refine") or 0.0) + (prof.get("global_rescan") or 0.0) + (prof.get("global_exact3") or 0.0)
        # the matching core = the tensor-core filter kernel + the exact refinement (refine + rescan); `achieved` charges both
        core_ms = (k_ms + r_ms) if k_ms else None
        achieved = (ALGO_FLOP_GLOBAL / (core_ms * 1e-3) / 1e12) if core_ms else None
        kname = ("gm_umma2_kernel (global matching, cta_group::2 tcgen05, three fp16-split products: every pair at fp32 grade)" if GM_EXACT3 else
                 "gm_fr_kernel (global matching, cta_group::2 tcgen05: ONE fp16 product filters candidates under a rigorous error bound; "
                 "gm_refine_kernel (+ gm_rescan_kernel) re-evaluates the survivors exactly in fp32; `achieved`/`frac` charge BOTH, i.e. use core_ms = kernel_ms + refine_ms)")
        roofline = {"kernel": kname,
                    "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": (achieved / peaks["tflops"]) if achieved else None, "traffic": traffic,
                    "achieved_executed": (EXEC_FLOP_GLOBAL / (k_ms * 1e-3) / 1e12) if k_ms else None,
                    "frac_executed": (EXEC_FLOP_GLOBAL / (k_ms * 1e-3) / 1e12 / peaks["tflops"]) if k_ms else None,
                    "kernel_ms": k_ms, "kernel_share_of_step": (core_ms / (serial_s * 1e3 / K)) if core_ms else None,
                    "refine_ms": r_ms, "refine_kernel_ms": prof.get("global_refine"), "rescan_kernel_ms": prof.get("global_rescan"),
                    "other_engine_chain_ms": prof.get("global_exact3"), "prepass_ms": prof.get("global_prepass"), "core_ms": core_ms,
                    "frac_filter_kernel_only": (ALGO_FLOP_GLOBAL / (k_ms * 1e-3) / 1e12 / peaks["tflops"]) if k_ms else None,
                    "timed_in": "single-stream pass (same K steps; in the two-stream headline pass event timings include queueing for SMs)",
                    "peak_source": peaks["source"],
                    "algorithmic_flop_per_launch": ALGO_FLOP_GLOBAL,
                    "executed_mma_flop_per_launch": EXEC_FLOP_GLOBAL,
                    "local": {"kernels": "lm_pool + lm_convert (pre-pass), lm_umma_kernel (tcgen05 banded GEMM + transform + bilinear cells + per-object min)"
                                         if os.environ.get("MANET_LM_ENGINE", "")[:1].lower() != "s" else
                                         "window_dist_kernel, upsample_mask_min_kernel (CUDA-core engine; pre-pass not timed)",
                              "bound": "hbm (north star); measured bound: instruction issue, see DESIGN.md section 4",
                              "algorithmic_bytes": ALGO_BYTES_LOCAL,
                              "main_kernel_ms": prof["local_main"], "prepass_ms": prof["local_prepass"],
                              "achieved_gbs": (ALGO_BYTES_LOCAL / ((prof["local_main"] + prof["local_prepass"]) * 1e-3) / 1e9)
                              if prof["local_main"] and prof["local_prepass"] else None,
                              "peak_gbs": peaks["hbm_gbs"],
                              # what bounds lm_umma_kernel in fact: warp instructions / (4 schedulers x SMs x SM clock), from the
                              # instruction count of the round's ncu capture (profiles/roofline_traffic.json)
                              "issue_bound_us": (lm_inst / (4.0 * 148 * 1.965e9) * 1e6) if lm_inst else None,
                              "issue_bound_note": "smsp__inst_executed.sum of lm_umma_kernel / (4 IPC x 148 SMs x 1.965 GHz); the kernel runs "
                                                  "one CTA per SM (201 KB shared memory) at ~48 % issue-active"}}
        line = {"metric": "matched frames/sec (global+local, 480p, 5 obj)", "value": world * K / total_s,
                "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": total_s * 1e3 / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (results are exact fp32 distances; candidates are filtered by an fp16 tensor-core GEMM with fp32 accumulate)",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "parallelism": f"{world} independent sequences (1 per GPU), no data-path collective",
                           "l2": "flushed between timed steps (256 MiB write outside the event pair)",
                           "streams": "local-matching branch forked onto a second stream, joined before the step's end event",
                           "reference_operands": "steady state of a propagation: the annotated frame's operands are kept between steps (see first_frame for the rebuild-every-step figure)",
                           "timing": "CUDA events per step on the launching stream, summed; max over ranks"},
                "e2e": {"value": world * K / e2e_stream_s, "unit": "frames/s", "h2d_bytes_per_step": sess.h2d_bytes_per_streamed_step,
                        "d2h_bytes_per_step": sess.d2h_bytes_per_step, "ms_per_step": e2e_stream_s * 1e3 / K,
                        "mode": "streaming propagation through the host-buffer C-ABI session (manet_session_submit_host/_wait with "
                                "MANET_STEP_STREAM), two slots: every step uploads ITS new inputs from pinned host memory -- the new frame's "
                                "embedding [C,H,W] fp32 and the previous frame's labels -- and downloads both result maps; the annotated "
                                "frame + scribble labels are uploaded once per sequence (they are constant along test.py:237-259) and the "
                                "previous frame's embedding is last step's current frame, already on the device",
                        "full_copy": {"value": world * K / e2e_s, "ms_per_step": e2e_s * 1e3 / K,
                                      "h2d_bytes_per_step": sess.h2d_bytes_per_step, "d2h_bytes_per_step": sess.d2h_bytes_per_step,
                                      "mode": "all three embeddings + both label maps re-uploaded every step (two-slot pipelined); "
                                              "PCIe-bound, this was r01's e2e.value",
                                      "sync_value": world * K / e2e_sync_s, "sync_ms_per_step": e2e_sync_s * 1e3 / K}},
                "single_stream": {"value": world * K / serial_s, "ms_per_step": serial_s * 1e3 / K},
                "host_binding": numa,
                "first_frame": {"value": world * K / nocache_s, "ms_per_step": nocache_s * 1e3 / K,
                                "note": "MANET_STEP_NO_REF_CACHE: the reference side of global matching (annotated frame: bucketing, tensor-core "
                                        "image, fp32 copy) rebuilt every step -- what the first frame of a propagation pays; `value` is the steady "
                                        "state of the loop (test.py:237-259: annotated frame and scribble constant), which converts only the new frame"},
                "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline,
                "wall_s_timed_region": wall_dev}
        if sharded:
            roofline["sharded_global_1080p"] = sharded     # the path that communicates: see SCALE per-N lines
        if propagation:
            line["propagation_50"] = propagation
        if session:
            line["session_8_rounds"] = session
        if intvos_fwd:
            line["intvos_forward"] = intvos_fwd
        if seghead:
            line["seghead"] = seghead
            line["frame_step_with_seghead"] = {"ms": total_s * 1e3 / K + seghead["ms"],
                                               "frames_per_s": 1e3 / (total_s * 1e3 / K + seghead["ms"]),
                                               "note": "matching step (value) + DynamicSegHead, the two device-timed parts of one propagation frame"}
        if world == 1 and os.environ.get("MANET_BENCH_CPU", "1") == "1":
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    sess.close()
    if world > 1:
        dist.destroy_process_group()


def seghead_leg(dev, iters=10):
    """SURVEY 8f-2, the step right after the matching path: DynamicSegHead (IntVOS.py:488-525) on the step's own
    outputs at the headline shape (480p embedding, 5 objects), fed by its parts (no repeat/cat, IntVOS.py:663-670).
    Random-init weights of the reference architecture; CUDA events, L2 flushed between iterations."""
    import torch
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    torch.manual_seed(0)
    head = DynamicSegHead().to(dev).eval()
    gen = torch.Generator().manual_seed(77)
    cur = (0.1 * torch.relu(torch.randn(C, H, W, generator=gen))).to(dev)
    gmap = torch.rand(1, H, W, N_IDS, 1, generator=gen).to(dev)
    lmap = torch.rand(1, H, W, N_IDS, 1, generator=gen).to(dev)
    prev = torch.randint(0, N_IDS, (H // 8 + 1, W // 8 + 1), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int().to(dev)
    ids = torch.arange(N_IDS, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for i in range(iters + 3):
        flush.fill_(i & 0xFF)
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = head.forward_parts(cur, gmap, lmap, prev, ids)
        t.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(s.elapsed_time(t))
    ms = sum(times) / len(times)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        out = head.forward_parts(cur, gmap, lmap, prev, ids)
    t.record()
    torch.cuda.synchronize()
    ms_warm = s.elapsed_time(t) / iters
    px = N_IDS * H * W
    flop_pw = 2.0 * px * ((C + 3) * 256 + 3 * 256 * 256 + 256)
    flop_dw = 2.0 * px * 49 * ((C + 3) + 3 * 256)
    res = {"workload": f"DynamicSegHead forward, input [{N_IDS},{C + 3},{H},{W}] assembled from its parts, random-init weights",
           "ms": ms, "ms_back_to_back": ms_warm, "l2": "ms: L2 flushed (256 MiB write) before every call; ms_back_to_back: consecutive calls",
           "kernels_per_call": 10, "checksum": float(out.sum()),
           "pointwise_gflop": flop_pw / 1e9, "depthwise_gflop": flop_dw / 1e9,
           "achieved_tflops_algorithmic": (flop_pw + flop_dw) / (ms * 1e-3) / 1e12,
           "numerics": "depthwise fp32 CUDA cores; 1x1 convs as 3 fp16-split tcgen05 products, fp32 accumulate"}
    if os.environ.get("MANET_BENCH_CPU", "1") == "1":
        from oracle import manet_oracle as O
        state = {k: v.detach().cpu() for k, v in head.state_dict().items()}
        x = O.seghead_features(cur.cpu(), gmap.cpu(), lmap.cpu(), prev.cpu(), ids.cpu())
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        with torch.no_grad():
            want = O.dynamic_seghead_forward(state, x)
        res["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        res["cpu_cores"] = os.cpu_count()
        res["max_abs_err_vs_oracle"] = float((out.cpu() - want).abs().max())
    return res


def intvos_forward_leg(dev, iters=5):
    """BASELINE config 2: full IntVOS forward on one synthetic 480p frame triple, 5 objects -- random-init DeepLabv3+ /
    ResNet-101 + semantic embedding (torch/cuDNN: the caller of the hot path) -> global + local matching ->
    DynamicSegHead, all through networks.deeplab.IntVOS.forward (IntVOS.py:556-575).  Two weight settings: the plain random
    initialisation in eval mode (what the config names; identity batch norms let activations grow with depth, so the
    embedding scale is arbitrary) and the same network with the embedding's last batch norm calibrated to unit variance on
    this input (the scale a trained network's BN keeps).  Reports the local-matching guard statistic G and the engine that
    served local matching on these ARCHITECTURE-generated embeddings."""
    import torch
    from cvpr2020_manet_b200.config import cfg
    from cvpr2020_manet_b200.networks import IntVOS as api
    from cvpr2020_manet_b200.networks.deeplab import IntVOS
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = False, D_LOCAL
    try:
        torch.manual_seed(0)
        model = IntVOS(cfg).to(dev).eval()
        gen = torch.Generator().manual_seed(0)
        x = torch.randn(3, 3, 480, 854, generator=gen).to(dev)
        blob = lambda: torch.randint(0, N_IDS, (480 // 32, 854 // 32 + 1), generator=gen).repeat_interleave(32, 0).repeat_interleave(32, 1)[:480, :854]
        ref_lab = blob().view(1, 1, 480, 854).float().to(dev)
        prev_lab = blob().view(1, 1, 480, 854).float().to(dev)
        n_obj = torch.tensor([N_IDS - 1])
        out = {}
        for setting in ("random_init_eval", "embedding_bn_calibrated"):
            if setting == "embedding_bn_calibrated":
                with torch.no_grad():
                    pre = model.embedding_conv(model.relu1(model.bn1(model.seperate_conv(model.feature_extracter(x)))))
                    model.bn2.running_mean.copy_(pre.mean(dim=(0, 2, 3)))
                    model.bn2.running_var.copy_(pre.var(dim=(0, 2, 3)))
            times, t_emb = [], []
            for i in range(iters + 2):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                with torch.no_grad():
                    ev[0].record()
                    emb = model.extract_feature(x)
                    ev[1].record()
                    res = model(x, ref_lab, prev_lab, seq_names=["s"], gt_ids=n_obj, k_nearest_neighbors=1)
                    ev[2].record()
                torch.cuda.synchronize()
                if i >= 2:
                    t_emb.append(ev[0].elapsed_time(ev[1])); times.append(ev[1].elapsed_time(ev[2]))
            pred = res["s"]
            stats = api.local_match_guard_stats(H, W, C, N_IDS, D_LOCAL, dev)
            ms, ms_emb = sum(times) / len(times), sum(t_emb) / len(t_emb)
            out[setting] = {"forward_ms": ms, "extract_feature_ms": ms_emb, "matching_and_head_ms": ms - ms_emb,
                            "frames_per_s": 1e3 / ms, "logits_shape": list(pred.shape), "logits_finite": bool(torch.isfinite(pred).all()),
                            "embedding_abs_max": float(emb.abs().max()), "embedding_sq_norm_mean": float((emb ** 2).sum(1).mean()),
                            "local_guard": stats}
        out["workload"] = ("IntVOS.forward: input [3,3,480,854] (reference, previous, current frame), 5 objects, no memories; "
                           "DeepLabv3+/ResNet-101 + semantic embedding on torch/cuDNN (library code, default TF32 conv policy of "
                           "torch), matching + DynamicSegHead on the sm_100a kernels; CUDA events, 5 iterations after 2 warm-ups")
        return out
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


def propagation_leg(dev, T=50):
    """BASELINE config 3: propagation over a synthetic 50-frame 480p sequence, 5 objects -- per frame global matching
    against the annotated frame + global-map memory, local matching against the previous frame + local-map memory, the
    dynamic head, bilinear upsample to 480x854 + argmax, labels fed to the next frame (test.py:237-259), all on the
    device through engine.propagate_sequence.  Embeddings are synthetic (the backbone is out of scope) and resident."""
    import torch
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    torch.manual_seed(0)
    head = DynamicSegHead().to(dev).eval()
    gen = torch.Generator().manual_seed(321)
    base = 0.1 * torch.relu(torch.randn(C, H, W, generator=gen))
    embs = torch.empty(T, C, H, W, device=dev)
    for t in range(T):
        embs[t] = (base + 0.01 * t * torch.randn(C, H, W, generator=gen)).to(dev)
    scr = torch.full((H, W), -1, dtype=torch.int32)
    for o in range(N_IDS):
        scr[10 + 15 * o, 20:150] = o
        scr[5 + 15 * o:25 + 15 * o, 30 + 25 * o] = o
    first = torch.randint(0, N_IDS, (H // 8 + 1, W // 8 + 1), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int()
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = True, D_LOCAL
    scr_d, first_d = scr.to(dev), first.to(dev)
    try:
        res = {}
        for rep in range(2):                                   # first pass warms workspaces / packs the weights
            # the per-sequence memories are created before the timed region (the reference creates them lazily inside the
            # first frame; 0.64 GB of fresh cudaMalloc + fill would otherwise land in the measurement)
            gm = {"bench": torch.ones((104, H, W, N_IDS, 1), dtype=torch.float32, device=dev)}
            lm = ({"bench": torch.zeros((104, 9, H, W, N_IDS, 1), dtype=torch.float32, device=dev)},
                  {"bench": torch.zeros((104, 9), dtype=torch.float32, device=dev)})
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            s.record()
            out, _ = engine.propagate_sequence(embs, range(1, T), 0, scr_d, first_d, N_IDS - 1, head, (480, 854),
                                               gm, lm, "bench", 1, D_LOCAL)
            e.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            res = {"frames": T - 1, "device_ms_per_frame": s.elapsed_time(e) / (T - 1), "wall_ms_per_frame": wall * 1e3 / (T - 1),
                   "frames_per_s": (T - 1) / wall}
        hist = torch.bincount(out[T - 1].flatten(), minlength=N_IDS).tolist()
        res.update({"workload": f"{T}-frame 480p propagation, 5 objects: matching + memories + DynamicSegHead + upsample/argmax per frame, "
                                "synthetic resident embeddings, random-init head", "last_frame_label_histogram": hist,
                    "timing": "CUDA events around the whole loop (device) and host wall clock including all launches"})
        return res
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


def session_leg(dev, T=50, rounds=8):
    """BASELINE config 4: an 8-round simulated interactive session on a synthetic 50-frame 480p sequence, 5 objects.
    Every round: synthetic scribbles on a random annotated frame (round 1 through rough_ROI, test.py:229-230), the
    interaction branch (IntVOS.int_seghead, IntVOS.py:683-764: local self-match merged into the global-map memory, local-map
    bookkeeping, and the interaction head of the reference's default configuration -- DynamicSegHead(in_dim=C+2),
    IntVOS.py:554 with config.py:52 -- fed by its parts; previous-round labels from round 2 on), its logits -> labels of the
    annotated frame (test.py:212-216), then propagation forwards and backwards over all frames (test.py:237-285) with
    interaction_num = 1..8 driving the local-map round selection."""
    import torch
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    torch.manual_seed(0)
    head = DynamicSegHead().to(dev).eval()
    inter_head = DynamicSegHead(in_dim=C + 2).to(dev).eval()
    gen = torch.Generator().manual_seed(99)
    base = 0.1 * torch.relu(torch.randn(C, H, W, generator=gen))
    embs = torch.empty(T, C, H, W, device=dev)
    for t in range(T):
        embs[t] = (base + 0.01 * t * torch.randn(C, H, W, generator=gen)).to(dev)
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = True, D_LOCAL
    size = (480, 854)
    try:
        gm, lm = {}, ({}, {})
        per_round, int_ms = [], []
        frames_done = 0
        labels = {}                                   # frame -> [1,Hf,Wf] int64 labels of the latest round
        n_obj = torch.tensor([N_IDS - 1])
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        for rnd in range(1, rounds + 1):
            ann = int(torch.randint(1, T - 1, (1,), generator=gen))
            scr = torch.full((1, 1, H, W), -1, dtype=torch.int32)
            for o in range(N_IDS):
                y = int(torch.randint(5, H - 5, (1,), generator=gen)); x = int(torch.randint(5, W - 60, (1,), generator=gen))
                scr[0, 0, y, x:x + 50] = o
                scr[0, 0, y - 4:y + 4, x + 10] = o
            scr = scr.to(dev)
            t0 = time.perf_counter()
            if rnd == 1:
                scr = engine.rough_ROI(scr)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res, lm = engine.int_seghead(ref_frame_embedding=embs[ann:ann + 1], ref_scribble_label=scr.float(),
                                         prev_round_label=None if rnd == 1 else labels[ann].view(1, 1, *size).float(),
                                         global_map_tmp_dic=gm, local_map_dics=lm, interaction_num=rnd, seq_names=["bench"],
                                         gt_ids=n_obj, frame_num=[ann], first_inter=(rnd == 1), inter_seghead=inter_head)
            labels[ann], first = engine.upsample_argmax(res["bench"], size)
            e1.record()
            fwd, _ = engine.propagate_sequence(embs, range(ann + 1, T), ann, scr[0, 0], first, N_IDS - 1, head, size, gm, lm, "bench", rnd,
                                               D_LOCAL)
            bwd, _ = engine.propagate_sequence(embs, range(ann - 1, -1, -1), ann, scr[0, 0], first, N_IDS - 1, head, size, gm, lm, "bench",
                                               rnd, D_LOCAL)
            labels.update(fwd); labels.update(bwd)
            torch.cuda.synchronize()
            per_round.append((time.perf_counter() - t0) * 1e3)
            int_ms.append(e0.elapsed_time(e1))
            frames_done += T - 1
        wall = time.perf_counter() - t_all
        dist = lm[1]["bench"][:T, :rounds]
        return {"workload": f"{rounds}-round session, {T} frames 480p, 5 objects: rough_ROI + interaction branch (matching, memories, "
                            "interaction head DynamicSegHead(in_dim=102) = the reference's default inter_seghead, labels) + bidirectional "
                            "propagation (matching, both memories, DynamicSegHead, labels) per round; interaction head included",
                "rounds": rounds, "propagated_frames": frames_done, "frames_per_s": frames_done / wall,
                "ms_per_round": [round(x, 2) for x in per_round], "interaction_branch_ms": [round(x, 3) for x in int_ms],
                "local_map_dist_table_nonzero": int((dist > 0).sum()), "global_map_min": float(gm["bench"][:T].min()),
                "last_round_label_histogram": torch.bincount(labels[0].flatten(), minlength=N_IDS).tolist()}
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


def sharded_1080p_leg(dev, rank, world, iters=5):
    """BASELINE config 5: large-reference global matching at 1080p shape (M = 129 600 queries, reference = T_mem stacked
    frames, T_mem in {1, 4, 8}) sharded over the reference axis; per-object partial minima combined with ONE
    all_reduce(MIN) of the [M,N] fp32 map (distributed.py).  Per T_mem: total ms (max over ranks),
    the matching part and the collective timed separately, per-GPU algorithmic TFLOP/s, and -- T_mem = 1 -- the parity of the
    all-reduced map against the unsharded result on the same tensors (at N=1: two shards emulated on the one GPU)."""
    import torch
    import torch.distributed as dist
    from cvpr2020_manet_b200.distributed import shard_bounds
    from cvpr2020_manet_b200.networks import IntVOS
    from cvpr2020_manet_b200 import memory
    Hb, Wb = 270, 480
    M = Hb * Wb
    peaks = load_peaks()
    gq = torch.Generator(device=dev).manual_seed(5)
    qry = (0.1 * torch.relu(torch.randn(C, Hb, Wb, generator=gq, device=dev))).permute(1, 2, 0)
    out = {}
    for t_mem in (1, 4, 8):
        R = t_mem * M
        b, e = shard_bounds(R, world, rank)
        rec = {"R": R}
        if t_mem == 1:
            # identical full reference on every rank (same seed, same generator), each rank matches its own slice
            gf = torch.Generator(device=dev).manual_seed(6)
            full = 0.1 * torch.relu(torch.randn(C, R, generator=gf, device=dev))
            full_lab = torch.randint(0, N_IDS, (R, 1, 1), generator=gf, device=dev).int()
            ref, lab = full[:, b:e].t().unsqueeze(1), full_lab[b:e]
            whole, _ = IntVOS.nearest_neighbor_features_per_object(full.t().unsqueeze(1), qry, full_lab, 1, N_IDS - 1)
            if world > 1:
                part, _ = IntVOS.nearest_neighbor_features_per_object(ref, qry, lab, 1, N_IDS - 1)
                dist.all_reduce(part, op=dist.ReduceOp.MIN)
                how = f"all_reduce(MIN) over {world} ranks vs the unsharded call on rank {rank}"
            else:
                h = R // 2
                p0, _ = IntVOS.nearest_neighbor_features_per_object(full[:, :h].t().unsqueeze(1), qry, full_lab[:h], 1, N_IDS - 1)
                p1, _ = IntVOS.nearest_neighbor_features_per_object(full[:, h:].t().unsqueeze(1), qry, full_lab[h:], 1, N_IDS - 1)
                part = torch.minimum(p0, p1)
                how = "two reference shards matched on the one GPU, torch.minimum, vs the unsharded call"
            err = ((part - whole).abs() / whole.abs().clamp_min(1.0)).max()
            nerr = (memory.normalize_distances(part) - memory.normalize_distances(whole)).abs().max()
            errs = torch.stack([err, nerr]).double()
            if world > 1:
                dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            rec["max_rel_err"], rec["max_abs_err_normalised"], rec["parity"] = float(errs[0]), float(errs[1]), how
            if float(errs[0]) > 2e-5 or float(errs[1]) > 1e-5:
                raise RuntimeError(f"sharded global matching differs from the unsharded result: {rec}")
            del full, whole
        else:
            g2 = torch.Generator(device=dev).manual_seed(100 + rank + 16 * t_mem)
            ref = (0.1 * torch.relu(torch.randn(C, e - b, generator=g2, device=dev))).t().unsqueeze(1)   # [R/G,1,C] view of [C,R/G]
            lab = torch.randint(0, N_IDS, (e - b, 1, 1), generator=g2, device=dev).int()
        t_all, t_match, t_red = [], [], []
        for i in range(iters + 2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            part, _ = IntVOS.nearest_neighbor_features_per_object(ref, qry, lab, 1, N_IDS - 1)
            ev[1].record()
            if world > 1:
                dist.all_reduce(part, op=dist.ReduceOp.MIN)
            ev[2].record()
            res = memory.normalize_distances(part)
            ev[3].record()
            torch.cuda.synchronize()
            if i >= 2:
                t_all.append(ev[0].elapsed_time(ev[3])); t_match.append(ev[0].elapsed_time(ev[1])); t_red.append(ev[1].elapsed_time(ev[2]))
        # the collective alone: ranks enter together, nothing to wait for
        red_alone = []
        if world > 1:
            buf = part.clone()
            for i in range(12):
                dist.barrier(); torch.cuda.synchronize()
                s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); dist.all_reduce(buf, op=dist.ReduceOp.MIN); t.record()
                torch.cuda.synchronize()
                if i >= 2:
                    red_alone.append(s.elapsed_time(t))
        ms = torch.tensor([sum(t_all) / len(t_all), sum(t_match) / len(t_match), sum(t_red) / len(t_red),
                           (sum(red_alone) / len(red_alone)) if red_alone else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        total_ms, match_ms, red_ms, red_alone_ms = (float(x) for x in ms)
        flop_rank = 2.0 * M * (R / world) * C
        tf = flop_rank / (match_ms * 1e-3) / 1e12
        rec.update({"n_gpus": world, "ms": total_ms, "match_ms": match_ms, "allreduce_in_step_us": red_ms * 1e3 if world > 1 else None,
                    "allreduce_us": red_alone_ms * 1e3 if world > 1 else None,
                    "allreduce_share_of_step": (red_alone_ms / total_ms) if world > 1 else 0.0,
                    "per_gpu_algorithmic_tflops": tf, "algorithmic_tflops": 2.0 * M * R * C / (total_ms * 1e-3) / 1e12,
                    "frac_of_peak": tf / peaks["tflops"],
                    "frac_of_sustained_peak": (tf / peaks["tflops_sustained"]) if peaks.get("tflops_sustained") else None})
        out[f"T_mem={t_mem}"] = rec
        del ref, lab, part, res
    out["workload"] = ("global matching, query 100x270x480 (M=129 600), reference T_mem stacked 1080p frames (R = T_mem*129 600), N=6, "
                       "reference axis sharded over the ranks; match_ms = scan+convert+tcgen05 GEMM+finalize of this rank's shard "
                       "(kernels of 3-56 ms: compare with the SUSTAINED peak), allreduce_us = the [M,N] fp32 all_reduce(MIN) timed alone")
    out["collective"] = "all_reduce(MIN) fp32 [129600,6] (3.1 MB), NCCL" if world > 1 else None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_own_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
