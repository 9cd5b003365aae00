"""Parser of the full-precision dumps written by oracle/_ref/eqtlbma_bf_ref_dump
(oracle/ref_dump_main.cpp): the UNMODIFIED reference's results at %.17g."""
from __future__ import annotations

import gzip


def parse_dump(path):
    op = gzip.open if str(path).endswith(".gz") else open
    out = {"subgroups": [], "genes": [], "pairs": [], "permjoin": {}, "permsep1": {}, "permsep2": {}, "end": None}
    cur = None
    with op(path, "rt") as fh:
        for line in fh:
            t = line.rstrip("\n").split("\t")
            k = t[0]
            if k == "SUBGROUPS":
                out["subgroups"] = t[1:]
            elif k == "GENE":
                out["genes"].append((t[1], int(t[2])))
            elif k == "PAIR":
                cur = {"gene": t[1], "snp": t[2], "nsub": int(t[3]), "ss": {}, "raw": {}, "w": {}}
                out["pairs"].append(cur)
            elif k == "SS":
                cur["ss"][int(t[1])] = (int(t[2]),) + tuple(float(x) for x in t[3:8])
            elif k == "RAW":
                cur["raw"][t[1]] = [float(x) for x in t[2:]]
            elif k == "W":
                cur["w"][t[1]] = float(t[2])
            elif k == "PERMJOIN":
                out["permjoin"][t[1]] = dict(nsnps=int(t[2]), pval=float(t[3]), nperm=int(t[4]),
                                             true=float(t[5]), med=float(t[6]), total=int(t[7]))
            elif k == "PERMSEP1":
                out["permsep1"][t[1]] = dict(nsnps=int(t[2]), pval=float(t[3]), nperm=int(t[4]),
                                             true=float(t[5]), total=int(t[6]))
            elif k == "PERMSEP2":
                out["permsep2"][(t[1], int(t[2]))] = dict(nsnps=int(t[3]), pval=float(t[4]), nperm=int(t[5]),
                                                          true=float(t[6]), total=int(t[7]))
            elif k == "END":
                out["end"] = (int(t[1]), int(t[2]))
    return out


def config_names(S, bfs):
    """names in gsl_combination order (gene_snp_pair.cpp:469-485)"""
    from itertools import combinations
    if bfs == "gen":
        return []
    names = []
    for k in range(1, S + 1):
        for comb in combinations(range(1, S + 1), k):
            names.append("-".join(str(c) for c in comb))
        if bfs == "sin":
            break
    return names
