#!/usr/bin/env python
"""bench.py -- site-pattern conditional-likelihood updates per second (BASELINE.json's metric).

Workload (N=1): BASELINE.json configs[1] -- primates.nex, 4-state GTR+G4, nruns=2 x nchains=4 = 8 MC^3
chains per analysis -- with R independent analyses ("replicas", --replicas, default 32) in flight on
ONE B200, the way the reference arm keeps one such analysis running on every host core.  A *step* is
one MCMC generation of every replica: one proposal per chain, the 8 chains of a replica evaluated in
ONE chain-batched engine call = one fused kernel launch (P(t) rebuild for the dirty branches, pruning
over the dirty nodes with the rescaler fused, root integration, lnL reduction), R launches per step
on R streams.  The site patterns, pattern weights and GTR eigensystem are the reference's own (taken
from the evaluation records in tests/golden, i.e. MrBayes' compressed matrix of primates.nex: 413
patterns, 898 sites); proposals are synthetic but MCMC-shaped: a branch move dirties one P(t) and the
path to the root, a parameter move (15 %) dirties the whole tree, 30 % of proposals are accepted,
rejects undo the index flips on the host.  `single_replica` in the output repeats the measurement
with one analysis alone on the GPU (the latency-bound regime of a single MrBayes run).

Legs, all on the same pre-generated cycle of steps, timed regions driven by a C host application
(mrbayes_b200/host/mb200_host_loop.c, a plain client of the C-ABI):
  value   device-resident replay (mb200_replay): job descriptors already in HBM, results left in HBM;
          per-step CUDA events on a control stream that fans out to / joins the replicas' streams;
          L2 flushed between steps (a 256 MB memset)
  e2e     the reference-facing C-ABI with HOST structs (mb200_evaluate_begin / _end per replica and
          generation: pack, launch with the job in the parameter block, 16-byte result records
          written by the kernel into pinned host memory); wall clock; a replica's generation g+1
          starts only after its generation g has returned its lnL values to the host;
          --host-threads host threads share the replicas
  roofline  algorithmic bytes of a step / device time of the step (value leg); avg_kernel_us is one
          launch alone, bracketed by events inside the engine
  cpu_baseline  the reference's own CPU kernels (oracle/_ref, FMA build) timed inside
          LaunchLogLikeForDivision on one host core, bounded sample

--impl reference times the reference's CPU path with every host core busy (N independent
serial `mb` processes; the reference has no threading and MPI is not installed).
Under torchrun (N>1) every rank drives its own GPU with its own R replicas (weak scaling:
independent runs never exchange state); one NCCL all-reduce of the per-run lnL sums (the
marginal-likelihood reduce of the reference's MPI build) closes the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "site-pattern CL updates/sec (node*pattern*rate)"
UNIT = "CL updates/s"
GOLD = ROOT / "tests" / "golden" / "primates_gtr_g4_fma.gold.gz"
REF_BIN = ROOT / "oracle" / "_ref" / "mb_b200"
REF_DATA = ROOT / "oracle" / "_ref" / "data" / "primates.nex"
BYTES_PER_UPDATE = {4: 51.0, 20: 243.0, 61: 744.0}      # SURVEY 8d (fp32, K=4 / 4 / 1)


# ------------------------------------------------------------------------------ workload
def primates_problem(n_chains: int, seed: int):
    """Problem with the reference's primates patterns/weights/model and one random tree per chain."""
    from mrbayes_b200 import abi, records, workloads
    divs, events = records.load(GOLD)
    d = divs[0]
    eig = next(e for e in events if e.kind in ("eigen",))
    first = next(e for e in events if e.kind == "eval").spec
    rng = np.random.default_rng(seed)
    trees = [workloads.random_tree(d.cfg["tip_count"], rng, mean_len=0.08) for _ in range(n_chains)]
    masks = np.stack([d.tips[t] for t in range(d.cfg["tip_count"])])
    pr = workloads.Problem(4, 4, d.cfg["pattern_count"], n_chains, trees, masks, d.weights[0], first.freqs,
                           eig.V, eig.Vinv, eig.lam, first.rates, 0.0, flags=abi.FLAG_NUC4_PINVAR_QUIRK)
    pr.allocate()
    return pr


def synthetic_problem(name: str, n_chains: int, seed: int):
    from mrbayes_b200 import workloads
    S, K, C, tips = {"nuc200k": (4, 4, 200_000, 32), "aa50k": (20, 4, 50_000, 64), "codon20k": (61, 1, 20_000, 32)}[name]
    return workloads.make_problem(S, K, C, tips, n_chains, seed=seed)


def make_cycle(pr, inst, n_steps: int, seed: int, p_full=0.15, p_accept=0.3):
    """Initial full evaluation, then a cycle of n_steps generations ending in a reset step that
    returns every chain to the post-initialisation state, so the cycle can be replayed forever."""
    from mrbayes_b200 import workloads
    rng = np.random.default_rng(seed)
    nch = pr.n_chains
    inst.evaluate([pr.full_evaluation(ch) for ch in range(nch)])
    snap = workloads.snapshot(pr)
    steps = []
    for _ in range(n_steps - 1):
        specs = []
        for ch in range(nch):
            old = pr.tree[ch].length.copy()
            if rng.random() < p_full:
                sp = pr.full_evaluation(ch)
            else:
                sp = pr.random_branch_update(ch, rng)
            specs.append(sp)
            if rng.random() >= p_accept:
                pr.reject(ch, sp, old)
        steps.append(specs)
    steps.append([workloads.reset_evaluation(pr, ch, snap) for ch in range(nch)])
    return steps


def updates_of(specs, C, K) -> int:
    return sum(len(s.ops) for s in specs) * C * K


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        try:
            sm = sorted(float(r[1]) for r in rows)
            reasons = set()
            for r in rows:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                    "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}
        except (ValueError, IndexError):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(rows)}


# ------------------------------------------------------------------------------ reference arm
def run_reference_process(ngen: int, seed: int, tmp: Path, tag: str):
    """One serial reference process in harness 'cpu' mode on the bench workload; -> Popen."""
    nex = tmp / f"{tag}.nex"
    nex.write_text(
        f"set autoclose=yes nowarn=yes seed={seed} swapseed={seed};\n"
        f"execute {REF_DATA};\nlset nst=6 rates=gamma ngammacat=4;\n"
        f"mcmc nruns=2 nchains=4 ngen={ngen} printfreq=1000000 samplefreq=1000000 diagnfreq=1000000 "
        f"filename={tmp / tag};\nquit;\n")
    env = dict(os.environ, MB200_MODE="cpu", MB200_REPORT=str(tmp / f"{tag}.json"))
    return subprocess.Popen([str(REF_BIN), str(nex)], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def reference_sample(n_procs: int, ngen: int, seed0: int):
    """n_procs concurrent serial reference processes; -> (sum of per-process kernel-only
    CL-updates/s, total CL updates, wall seconds, mean in-kernel seconds)."""
    with tempfile.TemporaryDirectory() as td:
        tmp = Path(td)
        t0 = time.perf_counter()
        procs = [run_reference_process(ngen, seed0 + i, tmp, f"p{i}") for i in range(n_procs)]
        for p in procs:
            p.wait()
        wall = time.perf_counter() - t0
        tot_rate, tot_upd, secs = 0.0, 0, []
        for i in range(n_procs):
            rep = json.loads((tmp / f"p{i}.json").read_text().strip().splitlines()[-1])
            tot_rate += rep["cl_updates"] / rep["sec_cpu"]
            tot_upd += rep["cl_updates"]
            secs.append(rep["sec_cpu"])
        return tot_rate, tot_upd, wall, float(np.mean(secs))


def usable_cores() -> int:
    """Host cores this process may actually use: the smaller of the CPU count, the scheduler affinity
    mask and the cgroup CPU quota (a container often sees every core of the machine but is capped)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path, parse in (("/sys/fs/cgroup/cpu.max", lambda t: (t.split()[0], t.split()[1])),):
        try:
            quota, period = parse(Path(path).read_text())
            if quota != "max":
                n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
        except (OSError, ValueError, IndexError):
            pass
    try:
        q = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
        per = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
        if q > 0 and per > 0:
            n = min(n, max(1, int(q / per + 0.5)))
    except (OSError, ValueError):
        pass
    return max(1, n)


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not REF_BIN.exists() or not REF_DATA.exists():
        # the oracle port is the documented stand-in when the reference binary did not travel
        val, sample = port_baseline(2.0)
        line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "value": val, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "primates.nex patterns, synthetic proposals",
                "config": {"workload": "primates.nex GTR+G4, oracle port (oracle/_ref missing)"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    cores = args.ref_procs if args.ref_procs > 0 else usable_cores()
    ngen = 1500                                     # ~0.7 s of in-kernel time per process
    for w in range(min(args.warmup, 1)):
        reference_sample(cores, 300, 900 + w)
    rates, upd, walls = [], 0, []
    steps = max(1, min(args.steps, 8))              # bounded: each step is ~1-2 s of all-core CPU work
    for s in range(steps):
        r, u, wall, _ = reference_sample(cores, ngen, 1000 + 100 * s)
        rates.append(r); upd += u; walls.append(wall)
    val = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "value": val, "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * float(np.mean(walls)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "examples/primates.nex (reference's own input), reference's own MCMC proposals",
            "config": {"workload": "primates.nex 4-state GTR+G4, nruns=2 nchains=4 (BASELINE configs[0]/[1])",
                       "reference_build": "unmodified sources, gcc -O3 -std=c99 -mavx -mfma (FMA kernels)",
                       "processes": cores, "ngen_per_process": ngen,
                       "timing": "time inside LaunchLogLikeForDivision (ld --wrap), summed rate over concurrent processes"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{cores} concurrent serial processes x {ngen} generations x {steps} steps"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def port_baseline(seconds: float):
    """Fallback CPU baseline: the oracle port replaying bench-shaped steps on one core."""
    from mrbayes_b200 import abi
    lib = abi.oracle_library()
    pr = primates_problem(8, 1)
    with pr.create(lib) as inst:
        inst.set_arith(1)
        steps = make_cycle(pr, inst, 64, 3)
        t0 = time.perf_counter(); upd = 0; n = 0
        while time.perf_counter() - t0 < seconds:
            sp = steps[n % len(steps)]
            inst.evaluate(sp); upd += updates_of(sp, pr.C, pr.K); n += 1
        dt = time.perf_counter() - t0
    return upd / dt, f"oracle port, {n} generations x 8 chains in {dt:.1f} s"


def cpu_baseline():
    if REF_BIN.exists() and REF_DATA.exists():
        rate, upd, wall, sec = reference_sample(1, 4000, 12345)
        return {"value": rate, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"unmodified reference (FMA kernels), primates GTR+G4 nruns=2 nchains=4, 4000 generations: "
                          f"{upd} CL updates in {sec:.2f} s inside LaunchLogLikeForDivision ({wall:.1f} s wall)"}
    val, sample = port_baseline(3.0)
    return {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}


# ------------------------------------------------------------------------------ engine arm
def other_workload(torch, lib, name, peaks, device):
    """Large synthetic configs (inputs >> L2): full-tree evaluations, resident, event-timed."""
    pr = synthetic_problem(name, 1, 2026)
    with pr.create(lib, device=device) as inst:
        stream = torch.cuda.ExternalStream(inst.stream(), device=device)
        inst.evaluate(pr.full_evaluation(0))
        batch = inst.pack([pr.full_evaluation(0)])
        inst.set_kernel_timing(True)
        for _ in range(2):
            inst.replay(batch)
        inst.kernel_time()
        reps = 5
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            inst.replay(batch)
        b.record(stream)
        inst.synchronize()
        ms = a.elapsed_time(b) / reps
        kms, kn = inst.kernel_time()
        upd = pr.n_int * pr.C * pr.K
        bpu = BYTES_PER_UPDATE[pr.S]
        flops = upd * (4 * pr.S * pr.S + pr.S)
        ach = upd * bpu / (kms / kn * 1e-3) / 1e9
        return {"workload": f"{name}: S={pr.S} K={pr.K} C={pr.C} taxa={pr.n_tips}, full-tree evaluation, 1 chain, working set > L2",
                "value": upd / (ms * 1e-3), "unit": UNIT, "ms_per_evaluation": ms,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": ach / peaks["hbm_gbs"], "traffic": None,
                             "tflops": flops / (kms / kn * 1e-3) / 1e12}}


def replica_seeds(rank: int, replicas: int):
    """(problem seed, proposal-cycle seed) of every replica of a rank: disjoint across ranks, so that
    under torchrun every GPU drives its own independent analyses (weak scaling)."""
    return [(20260924 + 1000 * rank + r, 7 + 1000 * rank + r) for r in range(replicas)]


def reduce_over_ranks(torch, dist, device, ms_value, ms_warm, ms_e2e, updates, launches):
    """The multi-rank contract of the bench line: times are the MAX over ranks, work is the SUM."""
    if dist is None:
        return ms_value, ms_warm, ms_e2e, float(updates), int(launches)
    vals = torch.tensor([ms_value, ms_warm, ms_e2e, float(updates), float(launches)], dtype=torch.float64, device=device)
    mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx[0].item(), mx[1].item(), mx[2].item(), sm[3].item(), int(sm[4].item())


def bench_engine(args):
    import torch
    from mrbayes_b200 import abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    peaks = {"hbm_gbs": 6650.0, "which": "fallback"}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = dict(json.loads(pk.read_text()), which="measured")

    lib = abi.engine_library()
    if lib.fn("device_count")() < 1:
        raise SystemExit("bench.py: no sm_100 device; the engine has no CPU fallback")

    n_chains = 8
    R = max(1, args.replicas)
    seeds = replica_seeds(rank, R)
    probs = [primates_problem(n_chains, seed=sd[0]) for sd in seeds]
    # several analyses share the GPU: SM time counts, not the latency of one launch
    inst_flags = abi.CONFIG_THROUGHPUT if (R > 1 and args.throughput_tiling) else 0
    insts = [p.create(lib, device=local, max_evaluations=n_chains, flags=inst_flags) for p in probs]
    pr, inst = probs[0], insts[0]
    cycle_len = 128
    steps_r = [make_cycle(p, i, cycle_len, seed=sd[1]) for sd, p, i in zip(seeds, probs, insts)]
    steps = steps_r[0]
    K, W = args.steps, args.warmup
    order = [i % cycle_len for i in range(K)]
    upd_per_step = [sum(updates_of(sr[i], pr.C, pr.K) for sr in steps_r) for i in range(cycle_len)]
    nodes_per_eval = float(np.mean([len(sp.ops) for sr in steps_r for s in sr for sp in s]))
    total_updates = sum(upd_per_step[i] for i in order)

    # device-resident job descriptors + host-side ctypes arrays, all built before timing
    batches_r = [[i.pack(s) for s in sr] for i, sr in zip(insts, steps_r)]
    host_arrays_r = [[abi.make_eval_array(s) for s in sr] for sr in steps_r]
    lnl = np.zeros(n_chains * R); st = np.zeros(n_chains * R, np.int32)
    p_lnl, p_st = lnl.ctypes.data_as(C.POINTER(C.c_double)), st.ctypes.data_as(C.POINTER(C.c_int))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    sync_all = (lambda: ([i.synchronize() for i in insts], torch.cuda.synchronize()))

    # the timed loops run in C (mrbayes_b200/host/mb200_host_loop.c), a plain client of the C-ABI
    hl = C.CDLL(str(abi.ENGINE_LIB.parent / "libmb200_hostloop.so"))
    hl.mb200_host_generation_loop.restype = C.c_double
    hl.mb200_host_replay_loop.restype = C.c_double
    # host threads of the end-to-end loop: the ranks of a node share the usable host cores
    HT = max(1, min(args.host_threads, R, max(1, usable_cores() // max(world, 1))))
    inst_ids = (C.c_int * R)(*[i.handle for i in insts])
    batch_ids = (C.c_int * (R * cycle_len))(*[b for br in batches_r for b in br])
    step_ptrs = (C.c_void_p * (R * cycle_len))(*[C.cast(a, C.c_void_p) for hr in host_arrays_r for a in hr])

    def c_order(seq):
        return (C.c_int * len(seq))(*seq), len(seq)

    def replay_loop(seq, flush_buf):
        arr, n = c_order(seq)
        ms = hl.mb200_host_replay_loop(inst_ids, C.c_int(R), batch_ids, C.c_int(cycle_len), arr, C.c_int(n),
                                       C.c_void_p(flush_buf.data_ptr() if flush_buf is not None else None),
                                       C.c_size_t(flush_buf.numel() if flush_buf is not None else 0))
        if ms < 0:
            raise RuntimeError(f"mb200_host_replay_loop failed with code {ms}")
        return ms

    def barrier():
        if dist is not None:
            dist.barrier()
        sync_all()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- warm-up: whole cycles so the state is back at the cycle start ------------------
    warm = max(W, 3)
    replay_loop([i % cycle_len for i in range(((warm + cycle_len - 1) // cycle_len) * cycle_len)], None)

    # ---- value: resident replay, L2 flushed between steps ------------------------------
    launches0 = sum(i.launch_count() for i in insts)
    barrier()
    t_clock0 = time.perf_counter()
    ms_value = replay_loop(order, flush)
    launches = sum(i.launch_count() for i in insts) - launches0
    run_lnl = torch.zeros(2, dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:      # final marginal-likelihood style reduce (1 double per run)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dist.all_reduce(run_lnl); b.record(); torch.cuda.synchronize()
        ms_value += a.elapsed_time(b)
    barrier()
    finish_cycle = (cycle_len - K % cycle_len) % cycle_len
    finish_seq = [(K + i) % cycle_len for i in range(finish_cycle)]
    if finish_seq:
        replay_loop(finish_seq, None)                   # untimed: return to the cycle start

    # ---- warm-L2 variant (no flush), informational -------------------------------------
    ms_warm = replay_loop(order, None)
    if finish_seq:
        replay_loop(finish_seq, None)

    # ---- single-replica latency figures: one 8-chain analysis alone on the GPU ----------
    single = None
    if R > 1:
        one_ids = (C.c_int * 1)(insts[0].handle)
        arr, n = c_order(order)
        ms1 = hl.mb200_host_replay_loop(one_ids, C.c_int(1), batch_ids, C.c_int(cycle_len), arr, C.c_int(n),
                                        C.c_void_p(flush.data_ptr()), C.c_size_t(flush.numel()))
        if finish_seq:
            arrf, nf = c_order(finish_seq)
            hl.mb200_host_replay_loop(one_ids, C.c_int(1), batch_ids, C.c_int(cycle_len), arrf, C.c_int(nf), C.c_void_p(None), C.c_size_t(0))
        sec1 = hl.mb200_host_generation_loop(one_ids, C.c_int(1), step_ptrs, C.c_int(cycle_len), C.c_int(n_chains), arr, C.c_int(n), p_lnl, p_st, C.c_int(HT))
        if finish_seq:
            hl.mb200_host_generation_loop(one_ids, C.c_int(1), step_ptrs, C.c_int(cycle_len), C.c_int(n_chains), arrf, C.c_int(nf), p_lnl, p_st, C.c_int(1))
        upd1 = sum(updates_of(steps[i], pr.C, pr.K) for i in order)
        single = {"replicas": 1, "value": upd1 / (ms1 * 1e-3), "ms_per_step": ms1 / K,
                  "e2e": upd1 / sec1, "e2e_ms_per_step": sec1 * 1e3 / K, "unit": UNIT,
                  "note": "one nruns=2 x nchains=4 analysis alone on the GPU: the latency-bound regime"}

    # ---- the same chains as ONE analysis (nruns = 2R, nchains = 4) on ONE instance: one launch per generation ----
    one_instance = None
    if R > 1 and rank == 0 and not args.no_one_instance:
        big = primates_problem(n_chains * R, seed=424242)
        with big.create(lib, device=local, max_evaluations=n_chains * R) as bi:
            bsteps = make_cycle(big, bi, 32, seed=99)
            bb = [bi.pack(sp) for sp in bsteps]
            ids1 = (C.c_int * 1)(bi.handle)
            barr = (C.c_int * len(bb))(*bb)
            seq = [i % len(bb) for i in range(max(64, min(K, 512)) // len(bb) * len(bb))]
            arr1, n1 = c_order(seq)
            hl.mb200_host_replay_loop(ids1, C.c_int(1), barr, C.c_int(len(bb)), arr1, C.c_int(n1), C.c_void_p(None), C.c_size_t(0))
            msb = hl.mb200_host_replay_loop(ids1, C.c_int(1), barr, C.c_int(len(bb)), arr1, C.c_int(n1),
                                            C.c_void_p(flush.data_ptr()), C.c_size_t(flush.numel()))
            updb = sum(updates_of(bsteps[i], big.C, big.K) for i in seq)
            one_instance = {"chains": n_chains * R, "value": updb / (msb * 1e-3), "ms_per_step": msb / n1, "unit": UNIT,
                            "note": f"nruns={2 * R} x nchains=4 of the same alignment as ONE analysis on one instance: "
                                    f"all chains of a generation in one launch (device-resident replay, L2 flushed)"}

    # ---- roofline of the fused kernel: events inside the engine, replica 0 alone, flushed ----
    inst.set_kernel_timing(True)
    stream = torch.cuda.ExternalStream(inst.stream(), device=local)
    kt_ms, kt_n, kt_updates = 0.0, 0, 0
    for chunk0 in range(0, min(K, 4096), 1024):
        sub = order[chunk0:chunk0 + 1024]
        with torch.cuda.stream(stream):
            for i in sub:
                flush.zero_()
                inst.replay(batches_r[0][i])
        ms, n = inst.kernel_time()
        kt_ms += ms; kt_n += n; kt_updates += sum(updates_of(steps[i], pr.C, pr.K) for i in sub)
    inst.set_kernel_timing(False)
    done = min(K, 4096)
    for i in range((cycle_len - done % cycle_len) % cycle_len):
        inst.replay(batches_r[0][(done + i) % cycle_len])
    inst.synchronize()

    # ---- e2e: the C-ABI call with host structs ------------------------------------------
    arr, n = c_order(list(range(cycle_len)))            # warm the host path, end at cycle start
    hl.mb200_host_generation_loop(inst_ids, C.c_int(R), step_ptrs, C.c_int(cycle_len), C.c_int(n_chains), arr, C.c_int(n), p_lnl, p_st, C.c_int(HT))
    barrier()
    arr, n = c_order(order)
    sec_e2e = hl.mb200_host_generation_loop(inst_ids, C.c_int(R), step_ptrs, C.c_int(cycle_len), C.c_int(n_chains), arr, C.c_int(n), p_lnl, p_st, C.c_int(HT))
    if sec_e2e < 0:
        raise RuntimeError(f"mb200_host_generation_loop failed with code {int(sec_e2e)}")
    sync_all()
    t_clock1 = time.perf_counter()
    barrier()
    if rank == 0:
        sampler.stop()

    # ---- reduce over ranks: MAX time, SUM work -------------------------------------------
    ms_value, ms_warm, ms_e2e, all_updates, all_launches = reduce_over_ranks(
        torch, dist, f"cuda:{local}", ms_value, ms_warm, sec_e2e * 1e3, total_updates, launches)

    h2d = float(np.mean([sum(pack_bytes(sr[i]).bytes for sr in steps_r) for i in range(cycle_len)]))
    nt_small = int(os.environ.get("MB200_NT_SMALL", "256"))
    tiles = -(-pr.C // (nt_small // 4))
    d2h = R * n_chains * 16 * (tiles if tiles <= 16 else 1)   # 16-byte result records written into mapped host memory
    if rank == 0:
        clocks = sampler.summary(t_clock0, t_clock1)
        if clocks.get("samples", 0) < 3:
            clocks["note"] = "timed region shorter than the 100 ms sampling period; nearest samples used"
        k_avg_s = (kt_ms / max(kt_n, 1)) * 1e-3
        # dominant (only) kernel: the fused pruning kernel, R launches per step running concurrently;
        # achieved = algorithmic bytes of a step / device time of the step (events, value leg)
        ach = (all_updates / world) * BYTES_PER_UPDATE[4] / (ms_value * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": all_updates / (ms_value * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": warm, "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "primates.nex site patterns/weights/GTR eigensystem from the reference's own run (tests/golden); synthetic MCMC-shaped proposals",
            "config": {"workload": f"primates.nex 4-state GTR+G4, nruns=2 x nchains=4 = 8 chains per analysis (BASELINE configs[1]), "
                                   f"{R} independent analyses (replicas) in flight per GPU, one chain-batched launch per analysis and generation; "
                                   f"the reference arm runs one such analysis per host core",
                       "replicas_per_gpu": R,
                       "patterns": pr.C, "rate_categories": pr.K, "states": pr.S, "taxa": pr.n_tips, "chains_per_gpu": n_chains * R,
                       "mean_dirty_nodes_per_evaluation": nodes_per_eval, "cycle_steps": cycle_len,
                       "l2": "flushed between timed steps (256 MB memset); warm-L2 figure in value_l2_warm",
                       "sharding": "independent runs per GPU, no data-path collective; one NCCL all-reduce of per-run lnL sums in the timed region (N>1)"},
            # without the flush kernel between steps the multi-stream fan-out is bounded by the host's enqueue
            # rate, not by the GPU: only meaningful for a single replica
            "value_l2_warm": (all_updates / (ms_warm * 1e-3)) if R == 1 else None,
            "e2e": {"value": all_updates / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K,
                    "api": "mb200_evaluate_begin / _end per analysis and generation (C-ABI, host structs in, 8 x lnL out), "
                           f"{HT} host thread(s), C generation loop (mrbayes_b200/host/mb200_host_loop.c); an analysis' results of generation g are on the host before its generation g+1 starts"},
            "gpu_launches": all_launches,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": (ach / peaks["hbm_gbs"]) if ach else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel, `ncu --set full`
                         # (profiles/r01_launches_bench_primates.md: 8 full-tree evaluations = 6.7 MB algorithmic, L2 warm)
                         "traffic": 70400, "traffic_unit": "bytes per launch (ncu, L2 warm: the 2.7 MB working set of an analysis lives in L2)",
                         "kernel": f"eval_nuc4_kernel<K=4,NT={nt_small},FUSE=true> (device-resident replay), {R} concurrent launches per step",
                         "avg_kernel_us": k_avg_s * 1e6, "launches_timed": kt_n,
                         "bytes_per_update": BYTES_PER_UPDATE[4], "peak_source": peaks["which"],
                         "note": "avg_kernel_us: one launch alone (events around it, launch latency included); each launch is latency-bound by construction "
                                 "(2.7 MB working set, ~36 dirty nodes x 413 patterns); the step is bounded by how many such launches the SMs hold"},
            "clocks": clocks,
        }
        if single is not None:
            line["single_replica"] = single
        if one_instance is not None:
            line["all_chains_one_instance"] = one_instance
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        if args.other and world == 1:
            line["other_workloads"] = [other_workload(torch, lib, n, peaks, local) for n in args.other.split(",")]
        print(json.dumps(line), flush=True)
    for i in insts:
        i.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


class pack_bytes:
    """Size of the packed job a step ships host->device (header + DevEval + rates/frequencies + branch
    list + node records).  Small jobs ride in the kernel parameter block, i.e. inside the launch."""

    def __init__(self, specs):
        a16 = lambda x: (x + 15) & ~15
        n_mat = sum(len(s.mats) for s in specs)
        n_op = sum(len(s.ops) for s in specs)
        n_dbl = sum(len(s.rates) + len(s.cat_weights) + len(s.freqs) for s in specs)
        self.bytes = a16(a16(a16(a16(16) + 96 * len(specs)) + 8 * n_dbl) + 16 * n_mat) + 48 * n_op


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4096)
    ap.add_argument("--warmup", type=int, default=128)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-procs", type=int, default=0,
                    help="--impl reference: concurrent serial reference processes (0 = every usable host core)")
    ap.add_argument("--host-threads", type=int, default=8,
                    help="host threads the end-to-end loop deals the replicas out to (the reference arm uses every host core)")
    ap.add_argument("--no-one-instance", action="store_true", help="skip the informational all-chains-on-one-instance leg")
    ap.add_argument("--throughput-tiling", action="store_true",
                    help="create the instances with MB200_CONFIG_THROUGHPUT (one CTA per evaluation walks all pattern "
                         "tiles); measured slower than the default tiling at 32 replicas of this workload: 86 vs 75 us/step")
    ap.add_argument("--replicas", type=int, default=32,
                    help="independent analyses (engine instances) in flight per GPU; 1 = a single analysis (latency regime)")
    ap.add_argument("--other", default="nuc200k,aa50k,codon20k",
                    help="comma list of extra large workloads reported under other_workloads ('' = none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_engine(args)


if __name__ == "__main__":
    main()
