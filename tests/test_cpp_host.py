"""The C++ host mirror (treepl_b200/host/pl_calc_b200.h) compiles against the C ABI with the system g++,
links the in-tree library, and fails loudly (exception, not a CPU result) when no GPU is present."""
import os
import subprocess
import sys
import tempfile

import pytest

from treepl_b200 import build as tbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cstring>
#include <vector>
#include "treepl_b200/host/pl_calc_b200.h"
int main() {
    // examples/test.tre flat vectors (SURVEY.md 8c)
    std::vector<int> parents = {-1,0,1,1,0,4,5,5,7,7,4,10,10}, cc = {2,2,0,0,2,2,0,2,0,0,2,0,0}, fr = {0,1,0,0,1,1,0,1,0,0,1,0,0};
    std::vector<std::vector<int> > kids = {{4,1},{3,2},{},{},{10,5},{7,6},{},{9,8},{},{},{12,11},{},{}};
    std::vector<double> c = {1000,2000,4000,4000,1000,2000,3000,1000,2000,2000,2000,3000,3000}, lf(13, 1.0);
    std::vector<double> mn = {10,0,0,0,3,0,0,0,0,0,3,0,0}, mx = {10,10,1e20,1e20,10,10,1e20,10,1e20,1e20,4,1e20,1e20};
    std::vector<double> pmn(13,-1.0), pmx(13,-1.0), sd = {10,3.5,0,0,6.7,3.5,0,1.5,0,0,3.7,0,0}, sr(13, 0.08), du(13, 1.0);
    pmn[10] = 3; pmx[10] = 4;
    pl_calc_b200 plp;
    plp.setup_starting_bits(&parents, &cc, &fr, &c, &lf, &mn, &mx, &sd, &sr, &du);
    plp.pen_min = &pmn; plp.pen_max = &pmx; plp.children_vec = &kids;
    std::vector<int> fp(26);
    int np = tpl_generate_param_order_vector(13, fr.data(), 0, nullptr, fp.data());
    std::vector<double> x;
    try {
        plp.set_freeparams(np, false, &fp, &x);
        double f = plp.calc_pl(x);
        std::vector<double> g(np);
        plp.calc_gradient(x, &g);
        printf("GPU f=%.6f np=%d g0=%.6f\n", f, np, g[0]);
        // device-resident batched optimiser: two copies of the problem at smoothing 0.1 (examples/test.cppr8s)
        plp.set_smoothing(0.1);
        const double f01 = plp.calc_pl(x);
        const int B = 2;
        std::vector<double> X(np * B), F(B), lo(np), up(np);
        for (int p = 0; p < np; p++) { X[p * B] = X[p * B + 1] = x[p]; lo[p] = 1e-40; up[p] = 1e15; }
        for (int i = 0; i < 13; i++) if (fp[13 + i] != -1) { lo[fp[13 + i]] = mn[i] == 0 ? 1e-40 : mn[i]; up[fp[13 + i]] = mx[i]; }
        std::vector<int> status(B);
        tpl_opt_result r = plp.optimize_batch(B, X.data(), F.data(), lo.data(), up.data(), TPL_GRAD_AD, nullptr, status.data());
        printf("OPT F=%.6f F0=%.6f status=%d steps=%d\n", F[0], f01, status[0], r.steps);
        // the same two problems through the CV entry: start columns of a one-column matrix, probes of rows 0 and np - 1
        tpl_opt_options oo;
        memset(&oo, 0, sizeof(oo));
        oo.method = 1;
        std::vector<double> X1(np), F2(B), probes(2 * B), Xm(np * B), Fm(B);
        for (int p = 0; p < np; p++) { X1[p] = x[p]; Xm[p * B] = Xm[p * B + 1] = x[p]; }
        plp.optimize_batch(B, Xm.data(), Fm.data(), lo.data(), up.data(), TPL_GRAD_AD, &oo);          // truncated Newton, expanded batch
        std::vector<int> col(B, 0), rows = {0, 0, np - 1, np - 1};
        plp.optimize_batch_cols(B, 1, X1.data(), col.data(), F2.data(), lo.data(), up.data(), TPL_GRAD_AD, &oo, 2, rows.data(), probes.data());
        printf("COLS dF=%.3e dP0=%.3e dP1=%.3e\n", F2[0] - Fm[0], probes[0] - Xm[0], probes[2] - Xm[(np - 1) * B]);
    } catch (const std::exception& ex) {
        printf("EXCEPTION %s\n", ex.what());
        return 3;
    }
    return 0;
}
'''


def _compile_and_run():
    lib = tbuild.build()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        exe = os.path.join(td, "t")
        open(src, "w").write(SRC)
        subprocess.check_call([cxx, "-std=c++11", "-O1", "-I", ROOT, src, "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)])
        return subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cpp_host_mirror_compiles_links_and_fails_loudly_without_gpu():
    if _has_gpu():
        pytest.skip("a CUDA device is present: covered by test_cpp_host_mirror_on_the_gpu")
    res = _compile_and_run()
    assert res.returncode == 3 and "EXCEPTION" in res.stdout and "no CUDA device" in res.stdout, res.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_on_the_gpu():
    """the same program on the B200: single-point calls through the mirror == the oracle, and the device-resident
    batched optimiser reached through the mirror improves the objective and reports a convergence status."""
    res = _compile_and_run()
    assert res.returncode == 0 and res.stdout.startswith("GPU f="), res.stdout
    import numpy as np
    from golden_util import Golden
    from oracle import plo
    flat = dict(Golden("test").flat)
    flat["lf"] = np.ones(13)
    flat["start_dates"] = np.array([10, 3.5, 0, 0, 6.7, 3.5, 0, 1.5, 0, 0, 3.7, 0, 0], dtype=np.float64)
    flat["start_rates"] = np.full(13, 0.08)
    flat["start_durations"] = np.ones(13)
    o = plo.OracleProblem(flat)
    x0 = o.set_freeparams(False)
    f, g = o.calc_pl_and_gradient(x0)
    first = res.stdout.splitlines()[0].split()
    assert abs(float(first[1].split("=")[1]) - f) <= 1e-6 and abs(float(first[3].split("=")[1]) - g[0]) <= 1e-5 * abs(g[0]), (res.stdout, f, g[0])
    opt = [ln for ln in res.stdout.splitlines() if ln.startswith("OPT ")][0].split()
    assert float(opt[1].split("=")[1]) < float(opt[2].split("=")[1]) and opt[3] in ("status=1", "status=2", "status=3"), res.stdout
    # the CV entry (start columns + probes) through the mirror: the same optimum, the probed rows of it
    cols = [ln for ln in res.stdout.splitlines() if ln.startswith("COLS ")][0].split()
    assert all(float(t.split("=")[1]) == 0.0 for t in cols[1:]), res.stdout
