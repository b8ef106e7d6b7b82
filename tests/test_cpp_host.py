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
    } catch (const std::exception& ex) {
        printf("EXCEPTION %s\n", ex.what());
        return 3;
    }
    return 0;
}
'''


def test_cpp_host_mirror_compiles_links_and_fails_loudly_without_gpu():
    lib = tbuild.build()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        exe = os.path.join(td, "t")
        open(src, "w").write(SRC)
        subprocess.check_call([cxx, "-std=c++11", "-O1", "-I", ROOT, src, "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)])
        res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert res.returncode == 0 and res.stdout.startswith("GPU f="), res.stdout
    else:
        assert res.returncode == 3 and "EXCEPTION" in res.stdout and "no CUDA device" in res.stdout, res.stdout
