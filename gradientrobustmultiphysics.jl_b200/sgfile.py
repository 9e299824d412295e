"""SimplexGrid 2.1 text files (`simplexgrid("assets/2d_grid_cookmembrane.sg")`, examples/Example202_LinearElasticity2D.jl:39;
the reader itself lives in ExtendableGrids).  Sections: DIMENSION, NODES (n dim, then coordinates), CELLS (n, then dim+1 node
numbers + region per line), FACES (boundary faces: dim node numbers + region), END.  Node numbers are 1-based.

A mesh read from such a file carries its own CellNodes: spaces whose dofs sit on the nodes only (H1P1, L2P0) get a dof map that
owes nothing to this package's face / edge enumeration -- the enumeration-free parity cases of SURVEY.md 8c."""
from __future__ import annotations

import numpy as np

from .grid import ExtendableGrid


def parse_sg(text: str):
    """-> dict(dim, coords[nnodes, dim], cellnodes[ncells, dim+1], cellregions, bfacenodes[nbfaces, dim], bfaceregions).
    Token based: writers differ in how they break lines (gWriteSG puts a node per line, ExtendableGrids a number per line)."""
    lines = [ln for ln in text.splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
    if not lines or not lines[0].strip().startswith("SimplexGrid"):
        raise ValueError("not a SimplexGrid file")
    tok = " ".join(lines[1:]).split()
    out, i = {}, 0

    def take(n):
        nonlocal i
        if i + n > len(tok):
            raise ValueError("unexpected end of file")
        v = tok[i:i + n]
        i += n
        return v
    while i < len(tok):
        key = take(1)[0]
        if key == "END":
            break
        if key == "DIMENSION":
            out["dim"] = int(take(1)[0])
        elif key == "NODES":
            n, d = (int(v) for v in take(2))
            out["coords"] = np.array([float(v) for v in take(n * d)], dtype=np.float64).reshape(n, d)
        elif key in ("CELLS", "FACES"):
            if "coords" not in out:
                raise ValueError("NODES must come first")
            n = int(take(1)[0])
            d = out["coords"].shape[1]
            w = (d + 1 if key == "CELLS" else d) + 1
            a = np.array([int(v) for v in take(n * w)], dtype=np.int64).reshape(n, w)
            if key == "CELLS":
                out["cellnodes"], out["cellregions"] = a[:, :-1].astype(np.int32), a[:, -1].astype(np.int32)
            elif n:
                out["bfacenodes"], out["bfaceregions"] = a[:, :-1].astype(np.int32), a[:, -1].astype(np.int32)
        else:
            raise ValueError(f"unknown section {key!r}")
    if "coords" not in out or "cellnodes" not in out:
        raise ValueError("NODES / CELLS section missing")
    if out.get("dim", out["coords"].shape[1]) != out["coords"].shape[1]:
        raise ValueError("dimension mismatch")
    return out


def simplexgrid(path_or_dict) -> ExtendableGrid:
    """simplexgrid(filename): grid with the file's own node and cell numbering"""
    d = path_or_dict if isinstance(path_or_dict, dict) else parse_sg(open(path_or_dict).read())
    return ExtendableGrid(d["coords"], d["cellnodes"], d.get("cellregions"), d.get("bfacenodes"), d.get("bfaceregions"))
