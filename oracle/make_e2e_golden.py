#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/e2e_reference.json: the END-TO-END results of the reference
binary itself (oracle/_ref/treePL = the unmodified /root/reference/src built by oracle/Makefile against the real
NLopt 2.4.2 / ADOL-C 2.6.3) on its own example control files, `seed = 1`, `nthreads = 1`.

    python oracle/make_e2e_golden.py [names...]

Per config: Langley-Fitch value, final PL value, the chi-square table of cv.out, the dated tree and the rate tree
(src/main.cpp:497-880).  The control files are copied to a scratch directory with absolute tree / output paths
(the reference writes its outputs next to the working directory; /root/reference is read-only).
Runs in THIS container only (needs /root/reference); the JSON it writes is what travels.
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
EX = os.path.join(REF, "examples")
BIN = os.path.join(HERE, "_ref", "treePL")
OUT = os.path.join(ROOT, "tests", "golden", "e2e_reference.json")

CONFIGS = ["test", "pl4203", "M1155", "vib", "clock", "big_test"]


def run_config(name, work):
    src = open(os.path.join(EX, name + ".cppr8s")).read().splitlines()
    out_tree = os.path.join(work, name + ".out")
    cv_out = os.path.join(work, name + ".cv")
    lines = []
    for ln in src:
        key = ln.split("=")[0].strip()
        if key == "treefile":
            ln = "treefile = " + os.path.join(EX, ln.split("=", 1)[1].strip())
        if key in ("outfile", "cvoutfile", "seed", "nthreads"):
            continue
        lines.append(ln)
    lines += ["outfile = " + out_tree, "cvoutfile = " + cv_out, "seed = 1", "nthreads = 1"]
    cfg = os.path.join(work, name + ".cppr8s")
    open(cfg, "w").write("\n".join(lines) + "\n")
    t0 = time.time()
    res = subprocess.run([BIN, cfg], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    wall = time.time() - t0
    log = res.stdout
    rec = {"config": name + ".cppr8s", "returncode": res.returncode, "wall_seconds": round(wall, 2)}
    m = re.findall(r"after opt calc: ([-+0-9.eE]+)", log)
    if m:
        rec["final_pl"] = float(m[-1])
        rec["after_opt_calc"] = [float(v) for v in m]
    m = re.findall(r"lf calc: ([-+0-9.eE]+)", log)
    if m:
        rec["lf"] = float(m[-1])
    m = re.findall(r"numparams:\s*(\d+)", log)
    if m:
        rec["numparams"] = [int(v) for v in m]
    if os.path.exists(cv_out):
        rec["cv_out"] = open(cv_out).read().splitlines()
        tab = []
        for ln in rec["cv_out"]:
            mm = re.match(r"chisq: \(([-+0-9.eE]+)\) ([-+0-9.eEnaif]+)", ln)
            if mm:
                tab.append([float(mm.group(1)), float(mm.group(2))])
        rec["chisq"] = tab
    if os.path.exists(out_tree):
        rec["dated_tree"] = open(out_tree).read().strip()
    if os.path.exists(out_tree + ".r8s"):
        rec["rate_tree"] = open(out_tree + ".r8s").read().strip()
    rec["log_tail"] = log.splitlines()[-6:]
    return rec


def main():
    names = sys.argv[1:] or CONFIGS
    if not os.path.exists(BIN):
        raise SystemExit("build the reference first: make -C oracle ref")
    old = {}
    if os.path.exists(OUT):
        old = json.load(open(OUT))
    with tempfile.TemporaryDirectory() as work:
        for nm in names:
            old[nm] = run_config(nm, work)
            print(nm, {k: old[nm].get(k) for k in ("returncode", "wall_seconds", "lf", "final_pl", "chisq")}, flush=True)
    old["_how"] = ("oracle/_ref/treePL (unmodified reference sources, g++ -O3 -fopenmp -std=c++11, NLopt 2.4.2, ADOL-C 2.6.3), "
                   "seed = 1, nthreads = 1; regenerate with python oracle/make_e2e_golden.py")
    json.dump(old, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
