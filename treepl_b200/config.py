"""`.cppr8s` control-file reader (the reference's main.cpp:88-270 grammar).

Line oriented `key = value` or bare `key`; split on '=', both sides trimmed; a key only
counts when the WHOLE token matches, so `#cv` and `[cv]` are simply unknown keys (that is
how "comments" work in the reference).  Lines of length <= 1 are skipped.
Only the keys the hot path and its CV/restart callers consume are interpreted; the rest
are kept verbatim in `.raw`.
"""
from __future__ import annotations

import os
import re
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
        return smoothing_grid(self.cvstart, self.cvstop, self.cvmultstep)


def smoothing_grid(cvstart, cvstop, cvmultstep):
    """CV grid as the while-loop of main.cpp:668,814 walks it (`while(curcv >= cvstop) ... curcv *= cvmultstep`) after
    the fix-ups of main.cpp:275-284: start/stop swapped when start < stop, a step > 1 inverted (in float, as the
    reference's `1/(float)cvmultstep`).  A step of exactly 1 or <= 0 never terminates / never descends in the reference;
    it is rejected here."""
    import numpy as np
    start, stop, mult = float(cvstart), float(cvstop), float(cvmultstep)
    if start < stop:
        start, stop = stop, start
    if mult > 1:
        mult = float(np.float32(1.0) / np.float32(mult))
    if not (0.0 < mult < 1.0):
        raise ValueError("cvmultstep = %r: the smoothing grid needs a step in (0, 1) (or > 1, which is inverted)" % cvmultstep)
    if stop <= 0.0:
        raise ValueError("cvstop = %r: the geometric smoothing grid needs a positive lower end" % cvstop)
    g = []
    cur = start
    while cur >= stop:
        g.append(cur)
        cur *= mult
    return g


def _tokens(val):
    """main.cpp tokenises mrca / min / max values on commas and blanks (Tokenize(..., ",     "))."""
    return [t for t in re.split(r"[,\s]+", val or "") if t]


def _atof(tok):
    """atof: the longest numeric prefix, 0 when there is none"""
    m = re.match(r"\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)", tok or "")
    return float(m.group(0)) if m else 0.0


def _atoi(tok):
    m = re.match(r"\s*[-+]?\d+", tok or "")
    return int(m.group(0)) if m else 0


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
            cfg.numsites = _atoi(val)
        elif key == "smooth":
            cfg.smooth = _atof(val)
        elif key == "mrca":
            parts = _tokens(val)
            cfg.mrcas[parts[0]] = parts[1:]
        elif key == "min":
            parts = _tokens(val)
            cfg.mins[parts[0]] = _atof(parts[1])
        elif key == "max":
            parts = _tokens(val)
            cfg.maxs[parts[0]] = _atof(parts[1])
        elif key == "cv":
            cfg.cv = True
        elif key == "randomcv":
            cfg.randomcv = True
            cfg.cv = True                        # main.cpp:237-239 sets both
        elif key == "cvstart":
            cfg.cvstart = _atof(val)
        elif key == "cvstop":
            cfg.cvstop = _atof(val)
        elif key == "cvmultstep":
            cfg.cvmultstep = _atof(val)
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
            cfg.seed = _atoi(val)
        elif key == "nthreads":
            cfg.nthreads = _atoi(val)
        elif key == "outfile":
            cfg.outfile = val
        elif key == "cvoutfile":
            cfg.cvoutfile = val
        elif key == "opt":
            cfg.opt = _atoi(val)
        elif key == "optad":
            cfg.optad = _atoi(val)
    return cfg


def parse_config(path):
    with open(path) as fh:
        return parse_config_text(fh.read(), basedir=os.path.dirname(os.path.abspath(path)))
