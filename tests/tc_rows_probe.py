"""GPU probe (not a test): whole-evaluation time of the tensor-core path as a function of the rows per tile.
usage: python tests/tc_rows_probe.py [aa50k|codon20k ...]   (spawns one process per setting: MB200_TC_ROWS is read once)"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def one(name):
    import torch
    import bench
    from mrbayes_b200 import abi, workloads
    S, K, Cn, tips = bench.SYNTH[name]
    nch = int(os.environ.get("PROBE_CHAINS", "1"))
    pr = workloads.make_problem(S, K, Cn, tips, nch, seed=2026)
    lib = abi.engine_library() if not os.environ.get("MB200_PROBE_LIB") else abi.Library(str(ROOT / "mrbayes_b200" / "lib" / os.environ["MB200_PROBE_LIB"]), "mb200_")
    with pr.create(lib) as inst:
        inst.evaluate(pr.full_evaluation(0))
        batch = inst.pack([pr.full_evaluation(c) for c in range(nch)])
        stream = torch.cuda.ExternalStream(inst.stream())
        for _ in range(3):
            inst.replay(batch)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(10):
            inst.replay(batch)
        b.record(stream)
        inst.synchronize()
        ms = a.elapsed_time(b) / 10
        upd = pr.n_int * pr.C * pr.K * nch
        print(json.dumps({"workload": name, "chains": nch, "rows": os.environ.get("MB200_TC_ROWS", "auto"), "kernel": "serial" if os.environ.get("MB200_TC_SERIAL") else "queue" if os.environ.get("MB200_TC_QUEUE") else "pipelined", "ms": ms,
                          "frac": upd * bench.bytes_per_update(S, K) / (ms * 1e-3) / 1e9 / 6569.6}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        names = sys.argv[1:] or ["codon20k", "aa50k"]
        for n in names:
            for serial in ("0", "1"):
                for rows in ("auto",):
                    env = dict(os.environ)
                    env.pop("MB200_TC_ROWS", None); env.pop("MB200_TC_SERIAL", None); env.pop("MB200_TC_QUEUE", None)
                    if rows != "auto":
                        env["MB200_TC_ROWS"] = rows
                    if serial == "1":
                        env["MB200_TC_SERIAL"] = "1"
                    if serial == "queue":
                        env["MB200_TC_QUEUE"] = "1"
                    subprocess.run(["timeout", "120", sys.executable, __file__, "--one", n], env=env, check=False)
