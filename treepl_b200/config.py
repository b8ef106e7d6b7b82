"""`.cppr8s` control-file reader (the reference's main.cpp:88-270 grammar).

Line oriented `key = value` or bare `key`; split on '=', both sides trimmed; a key only
counts when the WHOLE token matches, so `#cv` and `[cv]` are simply unknown keys (that is
how "comments" work in the reference).  Lines of length <= 1 are skipped.
Only the keys the hot path and its CV/restart callers consume are interpreted; the rest
are kept verbatim in `.raw`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field


@dataclass
class Config:
    treefile: str = ""
    numsites: int = 0
    smooth: float = 10.0                 # main.cpp:62
    mrcas: dict = field(default_factory=dict)
    mins: dict = field(default_factory=dict)
    maxs: dict = field(default_factory=dict)
    cv: bool = False
    randomcv: bool = False
    cvstart: float = 1000.0              # main.cpp:71-73
    cvstop: float = 0.1
    cvmultstep: float = 0.1
    log_pen: bool = False
    thorough: bool = False
    prime: bool = False
    collapse: bool = False
    set1: bool = False
    seed: int = -1
    nthreads: int = 1
    outfile: str = ""
    cvoutfile: str = "cv.out"
    opt: int = 0
    optad: int = 0
    raw: dict = field(default_factory=dict)
    basedir: str = "."

    def calibrations(self):
        """-> [(tips, min_or_None, max_or_None)] keyed like main.cpp's mrca_mins / mrca_maxs
        maps (iteration in name order, mins before maxs - the flattener does not depend on it)."""
        out = []
        for nm in sorted(set(self.mins) | set(self.maxs)):
            if nm not in self.mrcas:
                raise ValueError("problem with mrca %s (probably no mrca named this)" % nm)
            out.append((self.mrcas[nm], self.mins.get(nm), self.maxs.get(nm)))
        return out

    def tree_path(self):
        return self.treefile if os.path.isabs(self.treefile) else os.path.join(self.basedir, self.treefile)

    def smoothing_grid(self):
        """CV grid as the while-loop of main.cpp:668,814 walks it (after the swap/invert fix-ups of :275-284)."""
        start, stop, mult = self.cvstart, self.cvstop, self.cvmultstep
        if start < stop:
            start, stop = stop, start
        if mult > 1:
            mult = 1.0 / mult
        g = []
        cur = start
        while cur >= stop:
            g.append(cur)
            cur *= mult
        return g


def parse_config_text(text, basedir="."):
    cfg = Config(basedir=basedir)
    for line in text.splitlines():
        if len(line) <= 1:
            continue
        toks = [t.strip() for t in line.split("=") if t != ""]
        if not toks:
            continue
        key = toks[0]
        val = toks[1] if len(toks) > 1 else None
        cfg.raw[key] = val
        if key == "treefile":
            cfg.treefile = val
        elif key == "numsites":
            cfg.numsites = int(val)
        elif key == "smooth":
            cfg.smooth = float(val)
        elif key == "mrca":
            parts = val.split()
            cfg.mrcas[parts[0]] = parts[1:]
        elif key == "min":
            parts = val.split()
            cfg.mins[parts[0]] = float(parts[1])
        elif key == "max":
            parts = val.split()
            cfg.maxs[parts[0]] = float(parts[1])
        elif key == "cv":
            cfg.cv = True
        elif key == "randomcv":
            cfg.randomcv = True
        elif key == "cvstart":
            cfg.cvstart = float(val)
        elif key == "cvstop":
            cfg.cvstop = float(val)
        elif key == "cvmultstep":
            cfg.cvmultstep = float(val)
        elif key == "log_pen":
            cfg.log_pen = True
        elif key == "thorough":
            cfg.thorough = True
        elif key == "prime":
            cfg.prime = True
        elif key == "collapse":
            cfg.collapse = True
        elif key == "set1":
            cfg.set1 = True
        elif key == "seed":
            cfg.seed = int(val)
        elif key == "nthreads":
            cfg.nthreads = int(val)
        elif key == "outfile":
            cfg.outfile = val
        elif key == "cvoutfile":
            cfg.cvoutfile = val
        elif key == "opt":
            cfg.opt = int(val)
        elif key == "optad":
            cfg.optad = int(val)
    return cfg


def parse_config(path):
    with open(path) as fh:
        return parse_config_text(fh.read(), basedir=os.path.dirname(os.path.abspath(path)))
