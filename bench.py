#!/usr/bin/env python
"""bench.py — site-updates/s of the colour-sweep rejection-KMC hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one full sweep (every colour class once) over the synthetic random-spin lattice of
the workload.  `value` is measured with the lattice resident in HBM (CUDA events on the library's
stream); `e2e` goes through the public C ABI with HOST buffers (pinned H2D of the lattice, the
sweep, D2H of the lattice + counters inside the timed region, every step).  N > 1 runs under
torchrun, one rank per GPU, weak scaling: every rank owns one slab of the single-GPU size.
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
                                       T=0.5, pin=0.0, cpu_n=(96, 96, 96),
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
# BASELINE.json configs[1] is the single-GPU configuration the contract names for N = 1; the 3-D
# configurations (the north-star target is quoted on sc/26n Potts) are measured in the same run and
# reported under "other_workloads"
DEFAULT_WORKLOAD = "ising_sq4n_16384"
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
        return {"value": per * cores / wall / 1e9, "unit": "Gsites/s", "cores": cores, "kind": "reference",
                "sample": sample, "seconds": wall}
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
    return {"value": per * cores / wall / 1e9, "unit": "Gsites/s", "cores": cores, "kind": "port",
            "sample": sample, "seconds": wall}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # 200 sweeps of the default workload are 0.18 s of device time: long enough for the clock sampler
    # (one nvidia-smi sample per 20 ms) to see the timed region several times
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, nargs=3, default=None, help="override the per-GPU lattice (debug)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-segment", type=int, default=20,
                    help="sweeps per host round trip of the e2e leg (examples/ising/in.ising: `stats 10.0` = 20 sweeps)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_workloads leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    w = dict(WORKLOADS[args.workload])
    if args.size:
        w["n"] = tuple(args.size)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    C = NCOLORS[w["lattice"]]
    b_alg = 4 * (C + 1)
    config = {"workload": args.workload, "description": w["desc"], "app_style": w["model"], "lattice": w["lattice"],
              "per_gpu_cells": list(w["n"]), "Q": w["q"], "temperature": w["T"], "pinned_fraction": w["pin"],
              "spin_type": "int32", "rng": "philox4x32-10 keyed on (seed, site, sweep, colour)",
              "decomposition": "%d slab(s) along %s" % (world, "z" if DIM[w["lattice"]] == 3 else "y"),
              "l2": "lattice (%.2f GB/GPU) exceeds the 126 MB L2; no flush needed" % (nsites(w["lattice"], w["n"]) * 4 / 1e9),
              "B_alg_bytes_per_site_update": b_alg}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(w, args.warmup, args.steps)
        per_step_ms = r["seconds"] / args.steps * 1e3
        line = {"impl": "reference", "metric": "site_updates_per_sec", "value": r["value"], "unit": "Gsites/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
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
        s.randomize(12345, 1, 2 if wl["model"] == "ising" else wl["q"])
        if wl["pin"] > 0:
            s.pin_random(777, wl["pin"])
        s.sync()
        return s

    def barrier(s):
        s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(s, warmup, steps, sample_clocks):
        """W untimed sweeps, then K sweeps between CUDA events on the library's stream; max over ranks."""
        s.sweeps(0, warmup)
        barrier(s)
        l0 = s.launch_count()
        clocks = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            clocks.start()
        s.reset_counters()
        s.timer_start()
        for i in range(steps):
            s.sweeps(warmup + i, 1)
        ms = s.timer_stop()
        barrier(s)
        clk = clocks.stop() if sample_clocks and rank == 0 else None
        launches = s.launch_count() - l0
        nacc, natt = s.counters()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        cnt = torch.tensor([natt, nacc, launches], dtype=torch.float64, device="cuda")
        if world > 1:
            per_rank = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(per_rank, t)
            if rank == 0:
                print("per-rank ms/step: " + " ".join("%.3f" % (float(x.item()) / steps) for x in per_rank),
                      file=sys.stderr)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        natt_all, nacc_all, launches_all = [int(v) for v in cnt.tolist()]
        return float(t.item()), natt_all, nacc_all, launches_all, clk

    s = make(w)
    ms, natt_all, nacc_all, launches_all, clk = timed(s, args.warmup, args.steps, True)
    sweep_no = args.warmup + args.steps
    value = natt_all / (ms * 1e-3) / 1e9

    # ---- e2e: the public C ABI with HOST buffers, as the SPPARKS plugin drives it for one `run`
    # segment of --e2e-segment (20) sweeps (plugin/app_cuda_mixin.h iterate_rejection): pinned H2D of the lattice,
    # the sweeps, each followed by a D2H read of the step's metric (accept/attempt counters), then D2H
    # of the lattice for the host-side output.  All copies are inside the timed region (wall clock).
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
    e2e_value = s.nlocal * world * seg / float(te[0].item()) / 1e9
    strict_value = s.nlocal * world * args.e2e_steps / float(te[1].item()) / 1e9

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # dominant kernel: 2-D lattices run one colour-fused launch per sweep (all C colour classes,
        # algorithmic bytes 4*N*(C+1) per launch); 3-D lattices one launch per colour class (each reads
        # the slab once and writes its own colour: 4*N*(1+1/C) per launch).  Either way a launch's
        # algorithmic bytes = B_alg x the site-updates it performs.
        fused = DIM[w["lattice"]] == 2
        dom_launches = args.steps * (1 if fused else C)
        bytes_per_launch = float(b_alg) * (natt_all / world) / dom_launches
        avg_launch_ms = ms / dom_launches
        achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
        kernel_name = ("sweep_tile2d_kernel (colour-fused, one launch per sweep)" if fused
                       else "sweep_stream_kernel (one launch per colour class)")
        line = {"metric": "site_updates_per_sec", "value": value, "unit": "Gsites/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                "accept_ratio": nacc_all / max(natt_all, 1),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                             "kernel": kernel_name,
                             "avg_launch_ms": avg_launch_ms, "algorithmic_bytes_per_launch": bytes_per_launch,
                             "colour_fused_bound_bytes_per_site": 8},
                "e2e": {"value": e2e_value, "unit": "Gsites/s",
                        "h2d_bytes_per_step": 4 * s.nlocal * world / seg,
                        "d2h_bytes_per_step": (4 * s.nlocal * world) / seg + 16 * world,
                        "definition": "one run segment through the C ABI as the SPPARKS plugin issues it: pinned H2D "
                                      "of the lattice, %d sweeps each followed by a D2H read of the counters, D2H of "
                                      "the lattice; bytes are the segment's copies divided by its sweeps (segment length as "
                                      "in examples/ising/in.ising: one stats line per 20 sweeps)" % seg,
                        "strict_per_step": {"value": strict_value, "unit": "Gsites/s", "steps": args.e2e_steps,
                                            "h2d_bytes_per_step": 4 * s.nlocal * world,
                                            "d2h_bytes_per_step": 4 * s.nlocal * world + 16 * world,
                                            "definition": "H2D of the lattice, one sweep, D2H of the lattice + "
                                                          "counters around every single sweep (PCIe-bound)"}},
                "gpu_launches": launches_all, "clocks": clk}
        traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_path):
            tr = json.load(open(traffic_path)).get(args.workload)
            if tr:
                line["roofline"]["traffic"] = tr
        if world == 1 and not args.no_extra and args.workload == DEFAULT_WORKLOAD and not args.size:
            # the other BASELINE.json single-GPU workloads, short runs (5 timed sweeps after 3 warm-up)
            s.close()
            del host
            torch.cuda.empty_cache()
            others = []
            for name in sorted(WORKLOADS):
                if name == args.workload:
                    continue
                wl = WORKLOADS[name]
                so = make(wl)
                oms, onatt, onacc, _, _ = timed(so, 3, 5, False)
                ov = onatt / (oms * 1e-3) / 1e9
                ob = 4 * (NCOLORS[wl["lattice"]] + 1)
                others.append({"workload": name, "description": wl["desc"], "value": ov, "unit": "Gsites/s",
                               "ms_per_step": oms / 5, "B_alg": ob, "roofline_frac": ov * ob / peak,
                               "accept_ratio": onacc / max(onatt, 1)})
                so.close()
            line["other_workloads"] = others
        if world == 1 and not args.no_cpu:
            r = cpu_reference(w, 1, 3)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
