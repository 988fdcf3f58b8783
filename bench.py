#!/usr/bin/env python
"""bench.py — particle-steps/s of 2DTissue's per-timestep particle update on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU step (rank 0 only)

A "step" is one pass of the hot path (bin -> sort -> neighbour/force/align -> Euler -> seam re-entry ->
re-projection) over all resident particles.  Workload (config.workload):
    c4shard  (default) the per-GPU shard of BASELINE.json configs[3]: 2M particles per GPU (16M on 8 GPUs), refined
             ("high-resolution") ellipsoid chart, Euclidean neighbour cutoff, fp32 fast path, weak scaling
    c2       configs[1]: 10k particles, Euclidean cutoff        c3  configs[2]: 1M particles, table criterion
    c5       configs[4]: noise sweep (order-parameter phase diagram): every GPU runs independent 1M-particle replicas,
             each with its own noise amplitude eta and seed; no communication; phi(eta) and mean speed in the JSON line
One JSON line is printed by rank 0.  `value` = whole-job particle-steps/s with the state resident in HBM;
`e2e` = the same metric through t2d_step_host() with pinned HOST buffers (H2D + step + D2H every step).
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SURFACE_AREA = 451.3          # ellipsoid_x4 surface area in mesh units^2
# algorithmic bytes per particle-step, SURVEY.md §8(d)
B_ALG = {("f32", "table"): 144, ("f32", "euclid"): 192, ("f64", "table"): 208, ("f64", "euclid"): 280}
# per-kernel algorithmic bytes per particle: the same table split by stage.  The fused
# Euclid kernel is credited with the stages it replaces (K3 + K4/K5), the counting-sort scatter with K1 + K2;
# no credit is taken for the traffic the fusion saves.
B_KERNEL = {
    ("f32", "euclid"): {"step_fused": (24 + 16) + 68, "scan": 0, "scatter": 16 + (36 + 16), "comm_pack": 0, "exchange_unpack": 0},
    ("f64", "euclid"): {"step_fused": (36 + 24) + 104, "scan": 0, "scatter": 16 + (52 + 32), "comm_pack": 0, "exchange_unpack": 0},
    ("f32", "table"): {"neigh_table": 24, "wrap_project": 68, "scan": 0, "scatter": 16 + 36},
    ("f64", "table"): {"neigh_table": 36, "wrap_project": 104, "scan": 0, "scatter": 16 + 52},
}


def measured_traffic(workload, dtype, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t["%s/%s/%s" % (workload, dtype, kernel)]["bytes_per_launch"]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
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
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def workload_spec(args, world):
    w = args.workload
    if w == "c4shard":
        per = args.particles_per_gpu or 2_000_000
        # "high-resolution ellipsoid mesh": the chart is refined with the particle count so that the particles per
        # face stay at 7-13 (2M on 149k faces ... 16M on 2.4M faces).  With a fixed mesh the reference's own physics
        # (distance-weighted lift + 1/d force) blows up as soon as ~25 particles share a face: 4M particles on the
        # 149k-face chart lose particles after ~10 steps on ONE GPU too (Validation.cpp:66-72 would throw).
        refine = 2 + (0 if world <= 1 else (1 if world <= 4 else 2))
        return dict(name="c4shard: %d particles/GPU, ellipsoid chart refined %d levels, Euclidean cutoff" % (per, refine),
                    per_gpu=per, total=per * world, mode="euclid", refine=refine, dtype=args.dtype or "f32")
    if w == "c2":
        return dict(name="c2: 10k particles, ellipsoid_x4 chart, Euclidean cutoff", per_gpu=10_000, total=10_000 * world,
                    mode="euclid", refine=0, dtype=args.dtype or "f32")
    if w == "c3":
        # the stock 4725-vertex chart puts 211 particles on every vertex (table distance 0 -> the reference's d := 0.001 rule ->
        # speeds of 1e2-1e5 chart widths): the reference itself would throw.  On the chart refined 2 levels (74.9 k vertices,
        # ~13 particles per vertex) the run is stable; its table only exists as thresholded CSR rows built on the GPU
        # k = 0.01: particles that share a nearest vertex are at table distance 0, which the reference turns into d = 0.001
        # (ForceHelper.cpp:59-62), i.e. a force of 1000 k |u_i - u_j| per pair; with k = 1 the mean speed is 85 chart widths per
        # unit time (8 % of the chart per step) and a run is one rare close pair away from losing particles
        return dict(name="c3: 1M particles, ellipsoid chart refined 2 levels (74.9k vertices), hop-count vertex-distance table as "
                         "thresholded CSR rows built on the GPU, k = 0.01", per_gpu=1_000_000, total=1_000_000 * world, mode="table",
                    refine=2, dtype=args.dtype or "f32", k=0.01)
    if w == "c5":
        per = args.particles_per_gpu or 1_000_000
        return dict(name="c5: noise sweep, %d independent replica(s) of %d particles, one per GPU, eta = k/(R-1) (R > 1) or 0.25, "
                         "ellipsoid chart refined 2 levels, Euclidean cutoff" % (world, per),
                    per_gpu=per, total=per * world, mode="euclid", refine=2, dtype=args.dtype or "f32", replicas=True)
    raise SystemExit("unknown workload " + w)


def sigma_for(total):
    return float(np.sqrt(0.5 * SURFACE_AREA / (np.pi * total)))   # packing fraction 0.5


def load_chart(t2d, refine):
    chart = t2d.load_chart(os.path.join(ROOT, "tests", "golden", "ellipsoid_x4.t2dchart"))
    if refine:
        chart = t2d.refine_chart(chart, refine)
    return chart


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path, bounded sample
# ------------------------------------------------------------------------------------------------------
def run_port_same_config(spec, steps, warmup, seconds_budget=60.0):
    """The CPU comparator on OUR arm's configuration: the O(N k) OpenMP restatement of the reference's step
    (oracle/t2d_oracle.c, bit-identical to the compiled reference on every golden fixture) on the same chart, sigma,
    particle count and seeded state as the GPU arm, all host threads, as many steps as fit the time budget.
    Returns a cpu_baseline dict (kind "port")."""
    t2d = importlib.import_module("2dtissue_b200")
    from oracle import oraclebind
    mode = 1 if spec["mode"] == "euclid" else 0
    N = spec["per_gpu"] if spec.get("replicas") else spec["total"]
    same = True
    refine = spec["refine"]
    if mode == 0 and refine > 0:   # the oracle holds the reference's dense V x V table: not on a 75 k-vertex chart (5.6 GB as uint8)
        refine, same, N = 0, False, min(N, 50_000)
    chart = load_chart(t2d, refine)
    orc = oraclebind.Oracle(chart)
    if mode == 0:
        orc.set_table(orc.build_hop_table())
    sigma = sigma_for(N) if mode == 1 else 0.4166666666666667
    eta = 0.25 if spec.get("replicas") else 0.0
    uv, n = t2d.seed_particles(N, seed=1234)
    r3d, vid, _ = orc.get_r3d(uv)
    st = dict(uv=uv, n=n, vid=vid, r3d=r3d)
    cores = os.cpu_count() or 1
    t_start = time.perf_counter()
    done_w = 0
    for _ in range(max(1, warmup)):   # untimed; cut short if one step already eats a third of the budget
        st = orc.step(st["uv"], st["n"], st["vid"], st["r3d"], 0.1, spec.get("k", 1.0), sigma, 0.001, eta=eta, seed=1234, mode=mode,
                      threads=cores, step_index=done_w)
        done_w += 1
        if time.perf_counter() - t_start > seconds_budget / 3:
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(max(1, steps)):
        st = orc.step(st["uv"], st["n"], st["vid"], st["r3d"], 0.1, spec.get("k", 1.0), sigma, 0.001, eta=eta, seed=1234, mode=mode,
                      threads=cores, step_index=done_w + done)
        done += 1
        if time.perf_counter() - t_start > seconds_budget:
            break
    dt = time.perf_counter() - t0
    return {"value": N * done / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": (("same configuration as the GPU arm" if same else "NOT the GPU arm's chart: dense-table oracle on the unrefined chart") +
                       " (%s): %d particles, sigma %.6g, %d-face chart, seeded state; O(N k) cell-list "
                       "restatement of the reference's step (oracle/t2d_oracle.c, OpenMP, %d threads), %d warm-up + %d timed steps"
                       % (spec["mode"], N, sigma, len(chart["faces"]), cores, done_w, done)),
            "same_config": same, "fault": int(st["fault"])}


def run_compiled_reference(spec, seconds_budget=8.0):
    """The UNMODIFIED compiled reference (oracle/_ref).  It is O(N^2) in time and memory and single-threaded, so it cannot
    hold the bench configuration: 1500 particles on its own unrefined chart.  In Euclid mode the harness feeds the
    reference's ForceHelper / OrientationHelper classes a Euclidean dist_length (the reference has no Euclid mode, so this
    is not simulate_flight end to end); in table mode it is the reference's simulate_flight."""
    try:
        from oracle import refbind, oraclebind
        if not refbind.available():
            return {"unavailable": "oracle/_ref is not built"}
        t2d = importlib.import_module("2dtissue_b200")
        chart = load_chart(t2d, 0)
        mode = 1 if spec["mode"] == "euclid" else 0
        ref = refbind.Ref()
        ref.chart_import(chart)
        if mode == 0:
            ref.table_import(oraclebind.Oracle(chart).build_hop_table())
        Ns = 1500
        sigma = sigma_for(Ns) if mode == 1 else 0.4166666666666667
        uv, n = t2d.seed_particles(Ns, seed=1234)
        r3d, vid = ref.get_r3d(uv)
        st = dict(uv=uv, n=n, vid=vid, r3d=r3d)
        st = ref.step(st["uv"], st["n"], st["vid"], st["r3d"], 0.1, 1.0, sigma, 0.001, mode=mode)
        t0 = time.perf_counter()
        done = 0
        while done < 50:
            st = ref.step(st["uv"], st["n"], st["vid"], st["r3d"], 0.1, 1.0, sigma, 0.001, mode=mode)
            done += 1
            if time.perf_counter() - t0 > seconds_budget:
                break
        dt = time.perf_counter() - t0
        return {"value": Ns * done / dt, "unit": "particle-steps/s", "cores": 1, "kind": "reference", "same_config": False,
                "sample": ("compiled reference (oracle/_ref), single thread (it has none), %d particles x %d steps on the unrefined "
                           "ellipsoid_x4 chart, %s criterion%s; O(N^2) time and memory: cannot hold the bench configuration"
                           % (Ns, done, spec["mode"],
                              " (harness-built Euclidean dist_length into the reference's ForceHelper/OrientationHelper, not "
                              "simulate_flight: the reference has no Euclid mode)" if mode == 1 else ""))}
    except Exception as e:   # a report, never a reason to lose the line
        return {"unavailable": str(e)[:200]}


def main_reference(args, rank, world):
    if rank != 0:
        return
    spec = workload_spec(args, world)
    cb = run_port_same_config(spec, args.steps, args.warmup, seconds_budget=90.0)
    cb["reference_compiled"] = run_compiled_reference(spec)
    v = cb["value"]
    line = {"impl": "reference", "metric": "particle_steps_per_sec", "value": v, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": spec["name"], "what_ran": cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))


def transport_parity_check(t2d, dist, rank, world, local_rank):
    """Slab runs only, before the timed workload: a small fp64 problem is stepped (a) by `world` ranks over
    the NCCL transport and (b), on rank 0, by one single context; the merged NCCL result must equal (b) BIT FOR BIT (global ids
    travel with the particles, the exact path sums in ascending id).  So every scaling line this bench prints also proves that
    the transport it timed moves the right particles.  Returns a small dict for the JSON line."""
    chart = load_chart(t2d, 0)
    N, steps = 40000, 5
    uv, n = t2d.seed_particles(N, seed=41)
    sigma = sigma_for(N)
    kw = dict(v0=0.1, k=1.0, sigma=sigma, step_size=0.001, eta=0.03, seed=5, neigh_mode=t2d.NEIGH_EUCLID,
              precision=t2d.PRECISION_FP64, device=local_rank)
    c0 = t2d.Context(chart, capacity=N, **kw)
    c0.set_particles(uv, n)
    s0 = c0.download(("uv", "n", "vid", "r3d"))
    ref = None
    if rank == 0:
        c0.step(steps)
        ref = c0.download()
    c0.close()
    cuts = t2d.slab_cuts(s0["r3d"][:N], world)
    uid = [t2d.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = t2d.Context(chart, capacity=N, **kw)
    ctx.comm_init(rank, world, uid[0], cuts)
    p = t2d.partition_by_slab(s0, cuts, rank)
    ctx.set_state(p["uv"], p["n"], p["vid"], p["r3d"], ids=p["ids"])
    fault = ctx.step(steps)
    part = ctx.download()
    part["ids"] = ctx.download_ids()
    ctx.close()
    parts = [None] * world
    dist.all_gather_object(parts, part)
    res = None
    if rank == 0:
        out = t2d.merge_by_id(parts, N)
        bad = {k: int(np.sum(out[k] != ref[k])) for k in ("n", "vid", "face", "color", "uv", "rdot", "r3d")}
        res = {"ok": fault == 0 and sum(q["ids"].size for q in parts) == N and not any(bad.values()),
               "what": "fp64, %d particles, %d steps, %d NCCL slabs vs one context: bit-identical" % (N, steps, world),
               "owned": [int(q["ids"].size) for q in parts], "mismatches": bad}
        if not res["ok"]:
            sys.stderr.write("transport parity check FAILED: %s\n" % json.dumps(res))
    return res


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def main_ours(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: lib2dtissue_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t2d = importlib.import_module("2dtissue_b200")
    spec = workload_spec(args, world)
    chart = load_chart(t2d, spec["refine"])
    mode = t2d.NEIGH_EUCLID if spec["mode"] == "euclid" else t2d.NEIGH_TABLE
    prec = t2d.PRECISION_FP32 if spec["dtype"] == "f32" else t2d.PRECISION_FP64
    Nloc = spec["per_gpu"]
    replicas = bool(spec.get("replicas"))            # c5: every rank is its own closed system
    sigma = sigma_for(spec["per_gpu"] if replicas else spec["total"]) if mode == t2d.NEIGH_EUCLID else 0.4166666666666667
    slabs = world > 1 and not replicas
    if slabs and mode != t2d.NEIGH_EUCLID:
        raise SystemExit("multi-GPU slabs support the Euclidean criterion (workloads c4shard / c2)")
    parity = transport_parity_check(t2d, dist, rank, world, local_rank) if slabs else None
    cap = int(Nloc * 1.25) + 65536 if slabs else Nloc      # room for halo copies and migration imbalance
    kw = dict(v0=0.1, k=spec.get("k", 1.0), sigma=sigma, step_size=0.001, neigh_mode=mode, precision=prec, capacity=cap, device=local_rank,
              lift_mode=t2d.LIFT_BARYCENTRIC if args.lift == "barycentric" else t2d.LIFT_REFERENCE)
    if args.lift == "barycentric":
        spec["name"] += ", barycentric lift (extension, not the reference's semantics)"
    if mode == t2d.NEIGH_TABLE:
        kw["table_kind"] = t2d.TABLE_HOPS_FROM_MESH
    eta, rep_seed = 0.0, 1234
    if replicas:   # OrientationHelper.cpp:67-70: heading = int(mean angle + eta * 360 * (u - 0.5)), u from (seed, step, id)
        eta = rank / (world - 1) if world > 1 else 0.25
        rep_seed = 1234 + 1000 * rank
        kw.update(eta=eta, seed=rep_seed)
    ctx = t2d.Context(chart, **kw)
    if not slabs:
        uv, n = t2d.seed_particles(Nloc, seed=rep_seed)
        ctx.set_particles(uv, n)
    else:
        # ONE global particle set (spec["total"]), seeded identically on every rank; every rank projects it on its own
        # GPU (deterministic), so equal-count slab cuts along x agree everywhere without communication; each rank then
        # uploads only the slab it owns, with global ids
        Ntot = spec["total"]
        uv, n = t2d.seed_particles(Ntot, seed=1234)
        r3d, vid = np.zeros(3 * Ntot), np.zeros(Ntot, dtype=np.int32)
        for a in range(0, Ntot, cap):
            b = min(Ntot, a + cap)
            rr, vv, _ = ctx.get_r3d(np.concatenate([uv[a:b], uv[Ntot + a:Ntot + b]]))
            m = b - a
            for k in range(3):
                r3d[k * Ntot + a:k * Ntot + b] = rr[k * m:(k + 1) * m]
            vid[a:b] = vv
        cuts = t2d.slab_cuts(r3d[:Ntot], world)
        uid = [t2d.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0], cuts)
        part = t2d.partition_by_slab(dict(uv=uv, n=n, vid=vid, r3d=r3d), cuts, rank)
        ctx.set_state(part["uv"], part["n"], part["vid"], part["r3d"], ids=part["ids"])
        del uv, n, r3d, vid
        ctx.step(0)                                        # builds the halo (collective), no time step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ctx.step(1)
    barrier()
    ctx.reset_counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    fault = ctx.step(args.steps)            # K steps back to back on the library's stream, synchronised at the end
    dev_ms = ctx.last_step_ms               # CUDA events on that stream around exactly these K steps
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    timed_counters = ctx.counters()
    launches = timed_counters["kernel_launches"]
    if fault:
        sys.stderr.write("rank %d: simulation fault mask %d during the timed steps (1 lost, 2 non-finite, 4 wrap cap, "
                         "8 migration, 16 message overflow): the measurement is invalid\n" % (rank, fault))
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tf = torch.tensor([(fault >> b) & 1 for b in range(8)], dtype=torch.int32, device="cuda")   # OR over ranks, bit by bit
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fault = sum(int(v) << b for b, v in enumerate(tf.tolist()))
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = spec["total"] * args.steps / (dev_ms * 1e-3)

    # per-kernel CUDA-event breakdown (extra steps, outside the timed region)
    prof = {}
    for _ in range(5):
        for name, ms in ctx.profile_step():
            prof.setdefault(name, []).append(ms)
    prof = {k: statistics.median(v) for k, v in prof.items()}
    by_rank = None
    if slabs:   # what every rank saw (kernel_ms above is rank 0's): the step is as slow as the slowest slab
        row = {"rank": rank, "owned": int(ctx.owned_count()), "step_fused_ms": prof.get("step_fused"),
               "exchange_unpack_ms": prof.get("exchange_unpack")}
        by_rank = [None] * world
        dist.all_gather_object(by_rank, row)

    phase = None
    if replicas:   # the phase diagram: relax every replica further (untimed), then read the polar order parameter
        if args.relax_steps > 0:
            fault |= ctx.step(args.relax_steps)
        ob = ctx.observables()
        row = {"rank": rank, "eta": eta, "seed": rep_seed, "phi": float(ob["phi"]), "mean_speed": float(ob["mean_speed"]),
               "steps_run": int(ctx.step_index), "lost": int(ob["lost"]), "fault": int(fault)}
        phase = [row]
        if dist is not None:
            phase = [None] * world
            dist.all_gather_object(phase, row)

    if args.no_cpu_baseline:
        if rank == 0:
            emit(json.dumps({"profiler_run": True, "ms_per_step": ms_per_step, "kernel_ms": prof, "fault": fault,
                             **({"phase_diagram": phase} if phase else {})}))
        if dist is not None:
            dist.destroy_process_group()
        return
    # end-to-end through the public call with pinned HOST buffers in the reference's layouts: single GPU = the
    # literal drop-in t2d_step_host (upload -> step -> download); slabs = upload of the owned slab + step + download
    e2e_steps = max(2, min(args.steps, 10))
    pin = lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()
    if not slabs:
        s = ctx.download(("uv", "n", "vid", "r3d"))
        h_uv, h_n, h_vid, h_r3d = pin(s["uv"]), pin(s["n"]), pin(s["vid"]), pin(s["r3d"])
        h_rdot, h_col = pin(np.zeros(2 * Nloc)), pin(np.zeros(Nloc, dtype=np.int32))
        # t2d_step_host_uv: r_UV and n go up (r_3D / vertices_3D_active are functions of r_UV and are re-projected on the
        # device), everything the reference's step produces comes back
        ctx.step_host(h_uv, h_n, h_vid, h_r3d, h_rdot, h_col, reproject=True)
        barrier()
        te = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.step_host(h_uv, h_n, h_vid, h_r3d, h_rdot, h_col, reproject=True)
        barrier()
        e2e_s = time.perf_counter() - te
        lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        host32 = os.environ.get("T2D_HOST32")
        host32 = (host32 != "0") if host32 is not None else (os.cpu_count() or 1) // max(1, lws) >= 4
        if spec["dtype"] == "f32" and host32:
            # fp32 contexts: the caller's arrays are doubles in the reference's layouts, but the library narrows / widens them on
            # host threads (Engine::step_host32) and the PCIe bus carries floats: uv + heading up, uv, r3d, rdot + 3 int arrays down
            # (only when the rank has >= 4 host cores to itself; otherwise doubles travel and the device converts)
            h2d = Nloc * (8 + 4)
            d2h = Nloc * (8 + 12 + 8 + 4 + 4 + 4)
        else:
            h2d = Nloc * (16 + 4)
            d2h = Nloc * (16 + 4 + 4 + 24 + 16 + 4)
    else:
        pinz = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
        hb = dict(uv=pinz(2 * cap, torch.float64), n=pinz(cap, torch.int32), vid=pinz(cap, torch.int32),
                  r3d=pinz(3 * cap, torch.float64), rdot=pinz(2 * cap, torch.float64), color=pinz(cap, torch.int32),
                  ids=pinz(cap, torch.int32).view(np.uint32))
        No = ctx.download_into(hb)
        h2d = d2h = 0
        barrier()
        te = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.set_state_raw(No, hb["uv"], hb["n"], hb["vid"], hb["r3d"], hb["ids"])
            ctx.step(1)
            h2d += No * (16 + 4 + 4 + 24 + 4)
            No = ctx.download_into(hb)
            d2h += No * (16 + 4 + 4 + 24 + 16 + 4 + 4)
        barrier()
        e2e_s = time.perf_counter() - te
        tb = torch.tensor([h2d / e2e_steps, d2h / e2e_steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tb)
        h2d, d2h = float(tb[0].item()) / world, float(tb[1].item()) / world   # per-rank means; scaled by world below
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = spec["total"] * e2e_steps / float(t.item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        key = (spec["dtype"], spec["mode"])
        dom = max(prof, key=prof.get) if prof else None
        roof = None
        if dom:
            bk = B_KERNEL[key][dom]
            ach = bk * Nloc / (prof[dom] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": measured_traffic(args.workload, spec["dtype"], dom) if world == 1 and not args.particles_per_gpu else None,
                    "peak_source": peak_src, "kernel_ms": prof[dom], "alg_bytes_per_particle": bk,
                    "step_alg_bytes_per_particle": B_ALG[key],
                    "step_achieved_gbs": B_ALG[key] * Nloc / (ms_per_step * 1e-3) / 1e9,
                    "step_frac": B_ALG[key] * Nloc / (ms_per_step * 1e-3) / 1e9 / peak}
        cb = run_port_same_config(spec, 4, 1, seconds_budget=25.0)
        cb["reference_compiled"] = run_compiled_reference(spec)
        line = {"metric": "particle_steps_per_sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
                "config": {"workload": spec["name"], "particles_total": spec["total"], "sigma": sigma,
                           "mesh_V": int(len(chart["uv"])), "mesh_F": int(len(chart["faces"])),
                           "l2": "per-step working set (%d MB) exceeds the 126 MB L2; no explicit flush" %
                                 int(Nloc * 112 / 1e6 + 32),
                           "parallelism": ("replicas%d: independent contexts, one per GPU, no communication" % world) if replicas else
                                          ("slab%d: x-slabs of one %d-particle set, NCCL halo + migration exchange per step" %
                                           (world, spec["total"])) if world > 1 else "single",
                           "fault": fault, **({"phase_diagram": phase} if phase else {}),
                           **({"weak_scaling_note": "per-GPU particle count is fixed, but the chart is refined with the total "
                               "(2/3/3/4 levels at 1/2/4/8 GPUs) and sigma follows the total particle count, so the per-GPU work "
                               "is similar, not identical, along the curve"} if args.workload == "c4shard" else {})},
                "clocks": clocks, "gpu_launches": int(launches), "wall_ms_per_step": wall_ms / args.steps,
                "counters": {k: int(timed_counters[k]) for k in ("pairs_in_range", "max_row", "order_fallbacks", "ties_cutoff", "wraps", "locate_fallbacks")},
                **({"transport_parity": parity} if parity is not None else {}),
                "kernel_ms": prof, **({"by_rank": by_rank} if by_rank else {}), "roofline": roof,
                "cpu_baseline": cb,
                "e2e": {"value": e2e_val, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d * world),
                        "d2h_bytes_per_step": int(d2h * world), "steps": e2e_steps}}
        emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: everything libraries print there while the bench runs (NCCL's version
    banner, for one) is sent to stderr; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (line + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4shard", choices=["c4shard", "c2", "c3", "c5"])
    ap.add_argument("--relax-steps", type=int, default=500, help="c5: untimed steps after the timed region, before phi is read")
    ap.add_argument("--particles-per-gpu", type=int, default=0)
    ap.add_argument("--dtype", default=None, choices=[None, "f32", "f64"])
    ap.add_argument("--lift", default="reference", choices=["reference", "barycentric"],
                    help="UV->3-D lift: the reference's distance weights (default, the headline) or the barycentric extension")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline and e2e legs (profiler runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    quiet_stdout()
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
