#!/usr/bin/env python3
"""Dynamic constant-time check of the secret-key kernels (hard part 7) — the run-time companion of the
static SASS audit (tools/ct_audit.py).  Runs tools/ct_dynamic_driver.py three times under ncu with the secrets all-zero,
all-one and random (all public inputs identical) and compares, kernel launch by kernel launch,
  warp instructions executed, thread instructions executed (predication included), global / local / shared memory
  requests and sectors / wavefronts (bank conflicts and uncoalesced accesses included).
Every counter must be IDENTICAL across the three classes: no branch, predicate or address may depend on a secret.

usage: tools/ct_dynamic.py [--n 16384] [--json profiles/r02_ct_dynamic.json]      (needs a GPU and ncu)
"""
import argparse, collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
           "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
           "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
           "smsp__inst_executed_op_branch.sum"]
SHARED_WAVEFRONTS = "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
KERNELS = "regex:k_comb|k_x25519|k_expand_key|k_sign_nonce|k_sign_finish|k_sk_convert"


def capture(cls, n, workdir):
    log = os.path.join(workdir, f"ct_dynamic_{cls}.csv")
    cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", KERNELS, "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "tools", "ct_dynamic_driver.py"), cls, str(n)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 or not os.path.exists(log):
        raise RuntimeError("ncu failed: " + (res.stderr or res.stdout)[-800:])
    lines = open(log).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    per = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        per.setdefault((int(r["ID"]), name), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return [(name, m) for (_, name), m in per.items()]


def run(n=16384, workdir=None):
    workdir = workdir or os.path.join(ROOT, "gpurun_out")
    os.makedirs(workdir, exist_ok=True)
    caps = {cls: capture(cls, n, workdir) for cls in ("zeros", "ones", "random")}
    names = [k for k, _ in caps["zeros"]]
    assert all([k for k, _ in caps[c]] == names for c in caps), "different kernel sequences"
    out = {"operations_per_launch": n, "metrics": METRICS, "classes": list(caps), "launches": [], "identical": True,
           "how": "every counter must be identical for all-zero, all-one and random secrets.  One exception, reported per launch as 'noisy': the "
                  "shared-memory WAVEFRONT counter (l1tex__data_pipe_lsu_wavefronts_mem_shared).  It includes replays caused by arbitration with the "
                  "kernel's local-memory traffic in the same L1 pipeline and is not reproducible: it differs by ~5e-5 between launches on IDENTICAL "
                  "inputs (e.g. three k_comb<0> launches with all-zero secrets: 579576 / 579570 / 579629) in no consistent order of the classes.  "
                  "For that one counter the criterion is a relative spread below 1e-3; the shared-memory load / store INSTRUCTION counts, which "
                  "secret-dependent predication would change, are exact like everything else"}
    for i, name in enumerate(names):
        row = {"kernel": name, "counters": {}, "identical": True, "noisy": []}
        for m in METRICS:
            vals = [caps[c][i][1].get(m) for c in caps]
            if len(set(vals)) == 1:
                row["counters"][m] = vals[0]
                continue
            row["counters"][m] = dict(zip(caps, vals))
            if m == SHARED_WAVEFRONTS and max(vals) - min(vals) <= 1e-3 * max(vals):
                row["noisy"].append(m)
            else:
                row["identical"] = False
        out["identical"] &= row["identical"]
        out["launches"].append(row)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--json")
    a = ap.parse_args()
    res = run(a.n)
    for r in res["launches"]:
        print(("SAME " if r["identical"] else "DIFF ") + r["kernel"], {k.split("__")[1][:28]: v for k, v in r["counters"].items()} if not r["identical"] else
              int(r["counters"]["smsp__inst_executed.sum"]), ("noisy (differs between two runs on equal inputs): %s" % r["noisy"]) if r["noisy"] else "", flush=True)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
    sys.exit(0 if res["identical"] else 1)
