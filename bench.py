#!/usr/bin/env python
"""bench.py — site-updates/s of the colour-sweep rejection-KMC hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one full sweep (every colour class once) over the synthetic random-spin lattice of
the workload.  `value` is measured with the lattice resident in HBM (CUDA events on the library's
stream, median of --repeats blocks of K sweeps, each block restarted from the same random lattice
and W warm-up sweeps); `e2e` goes through the public C ABI with HOST buffers (pinned H2D of the
lattice, the sweeps, D2H of the lattice + counters inside the timed region).  N > 1 runs under
torchrun, one rank per GPU, weak scaling: every rank owns one slab of the single-GPU size; before
anything is timed the N-slab run is checked against the 1-GPU run (`mgpu_parity`).
`--impl reference` times the reference's own CPU implementation (oracle/_ref/spk_ref replicas on
all host cores; the C oracle port if the reference binary is absent) on a bounded sample.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs[1..4]; per-GPU lattice (weak scaling replicates it along the slab dimension)
WORKLOADS = {
    "potts_neighonly_sc26n_1024": dict(model="potts/neighonly", lattice="sc/26n", n=(1024, 1024, 1024), q=10000,
                                       T=0.5, pin=0.0, cpu_n=(96, 96, 96), coarsen_to=100,
                                       desc="configs[2]: 3D sc/26n 1024^3 potts/neighonly Q=10000, random start"),
    "ising_sq4n_16384": dict(model="ising", lattice="sq/4n", n=(16384, 16384, 1), q=2, T=1.1346, pin=0.0,
                             cpu_n=(2048, 2048, 1),
                             desc="configs[1]: 2D sq/4n 16384^2 Ising Metropolis near T_c"),
    "potts_pin_sc26n_1024": dict(model="potts/pin", lattice="sc/26n", n=(1024, 1024, 1024), q=10000, T=0.5,
                                 pin=0.01, cpu_n=(128, 128, 128),
                                 desc="configs[3] per-GPU share: 3D sc/26n potts/pin Q=10000, 1% pinned"),
    "potts_tri_32768": dict(model="potts", lattice="tri", n=(32768, 16384, 1), q=1000, T=0.5, pin=0.0,
                            cpu_n=(2048, 1024, 1),
                            desc="configs[4]: 2D tri 32768x16384 cells x 2 basis, potts Q=1000"),
}
# BASELINE.json's metric names no configuration, so N = 1 runs the largest single-GPU configuration:
# configs[2] (1.07e9 sites), which is also the one the north-star's 60 % target is quoted on.  The
# other single-GPU workloads are measured in the same run with full records under "workloads".
DEFAULT_WORKLOAD = "potts_neighonly_sc26n_1024"
NCOLORS = {"sq/4n": 2, "sq/8n": 4, "tri": 4, "sc/6n": 2, "sc/26n": 8}
NBASIS = {"sq/4n": 1, "sq/8n": 1, "tri": 2, "sc/6n": 1, "sc/26n": 1}
MAXNEIGH = {"sq/4n": 4, "sq/8n": 8, "tri": 6, "sc/6n": 6, "sc/26n": 26}
DIM = {"sq/4n": 2, "sq/8n": 2, "tri": 2, "sc/6n": 3, "sc/26n": 3}


def nsites(lattice, n):
    return n[0] * n[1] * (n[2] if DIM[lattice] == 3 else 1) * NBASIS[lattice]


def dt_sweep(w):
    if w["model"] == "ising":
        return 0.5
    if w["model"] == "potts/neighonly":
        return 1.0 / MAXNEIGH[w["lattice"]]
    return 1.0 / w["q"]


def make_config(name, w, world):
    C = NCOLORS[w["lattice"]]
    return {"workload": name, "description": w["desc"], "app_style": w["model"], "lattice": w["lattice"],
            "per_gpu_cells": list(w["n"]), "Q": w["q"], "temperature": w["T"], "pinned_fraction": w["pin"],
            "spin_type": "int32", "rng": "philox4x32-10 keyed on (seed, site, sweep, colour)",
            "decomposition": "%d slab(s) along %s" % (world, "z" if DIM[w["lattice"]] == 3 else "y"),
            "l2": ("int32 lattice %.2f GB/GPU; the sweeps run on its 1-bit copy (%.1f MB), which is the workload's whole working "
                   "set in that representation and stays in the 126 MB L2 between sweeps by design; no flush (re-reading it "
                   "from HBM every sweep would cost %.0f us of a %.0f us sweep)"
                   % (nsites(w["lattice"], w["n"]) * 4 / 1e9, nsites(w["lattice"], w["n"]) / 8 / 1e6,
                      nsites(w["lattice"], w["n"]) / 8 / 6.5e12 * 1e6, 257.0)
                   if w["model"] == "ising" and w["lattice"] == "sq/4n" and world == 1 else
                   "lattice (%.2f GB/GPU) exceeds the 126 MB L2; no flush needed" % (nsites(w["lattice"], w["n"]) * 4 / 1e9)),
            "B_alg_bytes_per_site_update": 4 * (C + 1)}


# ------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe), sampled during the timed region
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc, self.lines, self.device = None, [], device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the unmodified reference (oracle/_ref/spk_ref) on all host cores
def _ref_script(w, n, warm, steps):
    app = "ising" if w["model"] == "ising" else "%s %d" % (w["model"], w["q"])
    dim = DIM[w["lattice"]]
    zreg = "0 %d" % n[2] if dim == 3 else "-0.5 0.5"
    dt = dt_sweep(w)
    lines = ["seed 56789", "app_style " + app, "dimension %d" % dim, "lattice %s 1.0" % w["lattice"],
             "region box block 0 %d 0 %d %s" % (n[0], n[1], zreg), "create_box box", "create_sites box",
             "set site range 1 %d" % (2 if w["model"] == "ising" else w["q"])]
    if w["pin"] > 0:
        lines.append("pin %g 0 0" % w["pin"])
    lines += ["sweep color", "sector no", "temperature %r" % w["T"]]
    if warm > 0:
        lines.append("run %.17g" % (dt * (warm - 0.5)))
    lines.append("run %.17g" % (dt * (steps - 0.5)))
    return "\n".join(lines) + "\n"


def cpu_reference(w, warm, steps, cores=None):
    """Aggregate site-updates/s of `cores` concurrent serial replicas of the reference (no MPI in
    this image: SURVEY §8d fallback), each sweeping its own cpu_n lattice; = sum(attempts)/max(wall)."""
    cores = cores or len(os.sched_getaffinity(0))
    n = w["cpu_n"]
    per = nsites(w["lattice"], n) * steps
    spk = os.path.join(ROOT, "oracle", "_ref", "spk_ref")
    sample = "%d replicas x %s %s %s, %d warm-up + %d timed sweeps each" % (
        cores, w["model"], w["lattice"], "x".join(str(v) for v in n if v > 1), warm, steps)
    detail = {"replicas": cores, "per_replica_cells": list(n), "scaled_from_per_gpu_cells": list(w["n"]),
              "sweep": "color", "sector": "no", "warmup_sweeps": warm, "timed_sweeps": steps}
    if os.path.exists(spk):
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "in.bench")
            open(path, "w").write(_ref_script(w, n, warm, steps))
            procs = [subprocess.Popen([spk, "-in", path, "-log", "none", "-screen", os.path.join(tmp, "s%d" % i)],
                                      cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                     for i in range(cores)]
            for p in procs:
                p.wait()
            loops = []
            for i in range(cores):
                m = re.findall(r"Loop time of ([0-9.eE+-]+)", open(os.path.join(tmp, "s%d" % i)).read())
                loops.append(float(m[-1]))
        wall = max(loops)
        detail["rng"] = "Park-Miller ranapp of the reference (src/random_park.cpp), one stream per replica"
        return {"value": per * cores / wall / 1e9, "unit": "Gsites/s", "cores": cores, "kind": "reference",
                "sample": sample, "seconds": wall, "detail": detail}
    # C oracle port, one thread per core (ctypes releases the GIL)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_py as O
    walls = [0.0] * cores

    def work(i):
        o = O.Oracle(w["model"], w["lattice"], n[0], n[1], n[2], w["q"], w["T"])
        spins = np.random.default_rng(i).integers(1, (2 if w["model"] == "ising" else w["q"]) + 1,
                                                  size=o.nsites, dtype=np.int32)
        for s in range(warm):
            o.sweep_philox(spins, 1, s)
        t = time.perf_counter()
        for s in range(steps):
            o.sweep_philox(spins, 1, warm + s)
        walls[i] = time.perf_counter() - t

    ts = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    wall = max(walls)
    detail["rng"] = "philox4x32-10 (C oracle port)"
    return {"value": per * cores / wall / 1e9, "unit": "Gsites/s", "cores": cores, "kind": "port",
            "sample": sample, "seconds": wall, "detail": detail}


def _peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(name):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(name)
    return None


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--repeats", type=int, default=3, help="timed blocks of --steps sweeps; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, nargs=3, default=None, help="override the per-GPU lattice (debug)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-segment", type=int, default=20,
                    help="sweeps per host round trip of the e2e leg (examples/ising/in.ising: `stats 10.0` = 20 sweeps)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the other workloads")
    ap.add_argument("--no-mgpu-check", action="store_true", help="skip the N-GPU == 1-GPU check before timing")
    ap.add_argument("--no-plugin", action="store_true", help="skip the e2e_plugin leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    args.repeats = max(1, args.repeats)
    w = dict(WORKLOADS[args.workload])
    if args.size:
        w["n"] = tuple(args.size)
        w.pop("coarsen_to", None)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = make_config(args.workload, w, world)

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(w, args.warmup, args.steps)
        per_step_ms = r["seconds"] / args.steps * 1e3
        cb = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"impl": "reference", "metric": "site_updates_per_sec", "value": r["value"], "unit": "Gsites/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                # what this arm really ran: a bounded sample of the workload (`config` is the GPU arm's)
                "reference_sample": r["detail"],
                "e2e": {"value": r["value"], "unit": "Gsites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch  # first: its bundled NCCL must be the copy libspparks_cuda binds to
    import torch.distributed as dist
    from spparks_b200 import LatticeSweep

    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- N-GPU == 1-GPU, spin for spin, through both ghost-refresh paths, before anything is timed
    mgpu = None
    if world > 1 and not args.no_mgpu_check:
        from spparks_b200.selfcheck import mgpu_parity
        mgpu = mgpu_parity(dist, rank, world, local_rank, halo="p2p")
        if mgpu == "ok":
            mgpu = mgpu_parity(dist, rank, world, local_rank, halo="nccl")

    def make(wl):
        dim = DIM[wl["lattice"]]
        gn = list(wl["n"])
        gn[dim - 1] *= world  # weak scaling: one single-GPU lattice per rank, stacked along the slab dimension
        s = LatticeSweep(wl["model"], wl["lattice"], gn[0], gn[1], gn[2], nspins=wl["q"], temperature=wl["T"],
                         seed=12345, device=local_rank, rank=rank, nranks=world)
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.frombuffer(bytearray(LatticeSweep.comm_id()), dtype=torch.uint8).cuda()
            dist.broadcast(idt, 0)
            s.comm_init(bytes(idt.cpu().numpy().tobytes()))
        return s

    def reset(s, wl):
        s.randomize(12345, 1, 2 if wl["model"] == "ising" else wl["q"])
        if wl["pin"] > 0:
            s.pin_random(777, wl["pin"])
        s.sync()

    def barrier(s):
        s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def block(s, first_sweep, steps):
        """`steps` sweeps between CUDA events on the library's stream, a lap event after every sweep;
        returns (max over ranks of the block's ms, per-rank ms, this rank's per-sweep ms)."""
        barrier(s)
        s.timer_start()
        for i in range(steps):
            s.sweeps(first_sweep + i, 1)
            s.timer_lap()
        ms = s.timer_stop()
        barrier(s)
        laps = s.timer_laps()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        per_rank = [float(ms)]
        if world > 1:
            pr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(pr, t)
            per_rank = [float(x.item()) for x in pr]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), per_rank, laps

    def measure(name, wl, warmup, steps, repeats, min_block_s=0.0):
        """Full record of one workload: device-timed value (median of `repeats` blocks, each from the
        same random lattice + warm-up), clocks, roofline, e2e through the C ABI with host buffers.
        min_block_s > 0 lengthens the blocks of fast workloads so that the clock sampler (one sample
        per 20 ms) sees them several times."""
        C = NCOLORS[wl["lattice"]]
        b_alg = 4 * (C + 1)
        s = make(wl)
        if min_block_s > 0:
            reset(s, wl)
            s.sweeps(0, warmup)
            est, _, _ = block(s, warmup, 3)
            steps = max(steps, int(min_block_s / (est / 3 * 1e-3)) + 1)
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()  # before the first barrier: no rank forks between a barrier and its timer
        blocks, per_rank_all, laps_all = [], [], []
        natt = nacc = launches = 0
        for r in range(repeats):
            reset(s, wl)
            s.sweeps(0, warmup)
            s.sync()
            s.reset_counters()
            l0 = s.launch_count()
            ms, per_rank, laps = block(s, warmup, steps)
            launches = s.launch_count() - l0
            nacc, natt = s.counters()
            blocks.append(ms)
            per_rank_all.append(per_rank)
            laps_all.append(laps)
        clk = clocks.stop() if rank == 0 else None
        cnt = torch.tensor([natt, nacc, launches], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        natt_all, nacc_all, launches_all = [int(v) for v in cnt.tolist()]
        ms = statistics.median(blocks)
        value = natt_all / (ms * 1e-3) / 1e9
        sweep_no = warmup + steps
        rec = {"value": value, "unit": "Gsites/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
               "repeats_ms_per_step": [b / steps for b in blocks],
               "per_rank_ms_per_step": [[v / steps for v in pr] for pr in per_rank_all],
               "per_sweep_ms_rank0": {"min": min(laps_all[-1]), "median": statistics.median(laps_all[-1]),
                                      "max": max(laps_all[-1])},
               "accept_ratio": nacc_all / max(natt_all, 1), "gpu_launches": launches_all, "clocks": clk}
        kname, kper, kwidth = s.sweep_path()
        nv = s.neigh_variant(True)
        if nv[0]:
            rec["neigh_variant_at_end"] = {"variant": nv[0], "mean_distinct_neighbour_values": nv[1]}

        # ---- the same lattice further along the run (grain growth coarsens it: the candidate-list
        # kernels then work on a few distinct neighbour values instead of 26)
        if wl.get("coarsen_to") and wl["coarsen_to"] > sweep_no:
            s.sweeps(sweep_no, wl["coarsen_to"] - sweep_no)
            sweep_no = wl["coarsen_to"]
            s.reset_counters()
            cms, _, _ = block(s, sweep_no, steps)
            cacc, catt = s.counters()
            ct = torch.tensor([catt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ct)
            sweep_no += steps
            cv = float(ct.item()) / (cms * 1e-3) / 1e9
            nv = s.neigh_variant(True)
            rec["coarsened"] = {"after_sweeps": wl["coarsen_to"], "value": cv, "unit": "Gsites/s",
                                "ms_per_step": cms / steps, "accept_ratio": cacc / max(catt, 1),
                                "variant": nv[0], "mean_distinct_neighbour_values": nv[1]}

        # ---- e2e: the public C ABI with HOST buffers, as the SPPARKS plugin drives it for one `run`
        # segment of --e2e-segment sweeps (plugin/app_cuda_mixin.h iterate_rejection): pinned H2D of the
        # lattice, the sweeps, each followed by a D2H read of the step's metric (accept/attempt counters),
        # then D2H of the lattice for the host-side output.  All copies inside the timed region (wall clock).
        host = torch.empty(s.nlocal, dtype=torch.int32).pin_memory()
        s.download_ptr(host.data_ptr(), s.nlocal)
        barrier(s)
        t0 = time.perf_counter()
        s.upload_ptr(host.data_ptr(), s.nlocal)
        seg = max(1, args.e2e_segment)
        for _ in range(seg):
            s.sweeps(sweep_no, 1)
            sweep_no += 1
            s.counters()
        s.download_ptr(host.data_ptr(), s.nlocal)
        barrier(s)
        e2e_s = time.perf_counter() - t0
        # strictest reading, reported next to it: a full host round trip of the lattice around EVERY sweep
        barrier(s)
        t1 = time.perf_counter()
        for _ in range(args.e2e_steps):
            s.upload_ptr(host.data_ptr(), s.nlocal)
            s.sweeps(sweep_no, 1)
            sweep_no += 1
            s.download_ptr(host.data_ptr(), s.nlocal)
            s.counters()
        barrier(s)
        strict_s = time.perf_counter() - t1
        te = torch.tensor([e2e_s, strict_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        nl = s.nlocal
        rec["e2e"] = {"value": nl * world * seg / float(te[0].item()) / 1e9, "unit": "Gsites/s",
                      "h2d_bytes_per_step": 4 * nl * world / seg,
                      "d2h_bytes_per_step": (4 * nl * world) / seg + 16 * world,
                      "definition": "one run segment through the C ABI as the SPPARKS plugin issues it: pinned H2D "
                                    "of the lattice, %d sweeps each followed by a D2H read of the counters, D2H of "
                                    "the lattice; bytes are the segment's copies divided by its sweeps" % seg,
                      "strict_per_step": {"value": nl * world * args.e2e_steps / float(te[1].item()) / 1e9,
                                          "unit": "Gsites/s", "steps": args.e2e_steps,
                                          "h2d_bytes_per_step": 4 * nl * world,
                                          "d2h_bytes_per_step": 4 * nl * world + 16 * world,
                                          "definition": "H2D of the lattice, one sweep, D2H of the lattice + "
                                                        "counters around every single sweep (PCIe-bound)"}}
        del host
        s.close()
        torch.cuda.empty_cache()

        # ---- roofline of the dominant kernel: 2-D lattices run one colour-fused launch per sweep (all C
        # colour classes, algorithmic bytes 4*N*(C+1) per launch); 3-D lattices one launch per colour
        # class (each reads the slab once and writes its own colour: 4*N*(1+1/C) per launch).  Either way
        # a launch's algorithmic bytes = B_alg x the site-updates it performs.
        peak, peak_src = _peak()
        dom_launches = steps * kper
        bytes_per_launch = float(b_alg) * (natt_all / world) / dom_launches
        avg_launch_ms = ms / dom_launches
        achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
        rec["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "traffic": _traffic(name), "peak_source": peak_src,
                           "kernel": "%s (%d launch%s per sweep)" % (kname, kper, "" if kper == 1 else "es"),
                           "avg_launch_ms": avg_launch_ms, "algorithmic_bytes_per_launch": bytes_per_launch,
                           "colour_fused_bound_bytes_per_site": 8,
                           "device_spin_bits": 8 * kwidth if kwidth else 1}
        if kwidth != 4:
            rec["roofline"]["note"] = (
                "`achieved` and `frac` use the contract's algorithmic bytes (int32 spins, SURVEY 8d); the kernel sweeps "
                "a %s copy of the lattice kept next to the int32 one, so its real DRAM bytes (`traffic`) are lower "
                "and `frac` can exceed what an int32 kernel could reach" % ("16-bit" if kwidth == 2 else "1-bit"))
        if "coarsened" in rec:
            rec["coarsened"]["roofline_frac"] = rec["coarsened"]["value"] * b_alg / world / peak
        return rec

    rec = measure(args.workload, w, args.warmup, args.steps, args.repeats)

    others = {}
    if world == 1 and not args.no_extra and args.workload == DEFAULT_WORKLOAD and not args.size:
        for name in sorted(WORKLOADS):
            if name != args.workload:
                others[name] = measure(name, dict(WORKLOADS[name]), args.warmup, max(args.steps, 20), args.repeats, 0.12)
                others[name]["config"] = make_config(name, WORKLOADS[name], world)

    if rank == 0:
        line = {"metric": "site_updates_per_sec", "value": rec["value"], "unit": "Gsites/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": config}
        for k, v in rec.items():
            if k not in line:
                line[k] = v
        if world > 1:
            line["mgpu_parity"] = mgpu if mgpu is not None else "skipped"
            pr = rec["per_rank_ms_per_step"]
            print("per-rank ms/step by repeat: " + " | ".join(" ".join("%.3f" % v for v in r) for r in pr), file=sys.stderr)
        if world == 1 and not args.no_cpu:
            r = cpu_reference(w, 1, 3)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            for name, orec in others.items():
                r = cpu_reference(WORKLOADS[name], 1, 3)
                orec["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        if world == 1 and not args.no_plugin and not args.size:
            try:
                from tools.plugin_e2e import plugin_e2e
                line["e2e_plugin"] = plugin_e2e()
            except Exception as e:  # the plugin binary is built from /root/reference, absent there = not built
                line["e2e_plugin"] = {"unavailable": str(e)[:200]}
        if others:
            line["workloads"] = others
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
