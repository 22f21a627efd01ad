#!/usr/bin/env python3
"""How much of the benchmark surface could differ between the classic 256-case table (in-repo) and Lewiner's 33-case
tables (scikit-image, the reference's dependency): census of the cube cases of the active cells of the bench field.
A cell's topology can only differ where its case is AMBIGUOUS: a face with diagonal corners of equal sign
(face-ambiguous) or -- Lewiner's interior tests -- the body-diagonal configurations.  GPU tool:
    python tools/ambiguous_census.py [R]   -> one JSON line (goes into DESIGN.md section 2)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench import baked_triplane  # noqa: E402
from compare_skimage import face_ambiguous_cases  # noqa: E402
from sculptmate_b200 import runtime  # noqa: E402
from sculptmate_b200.tsr import TSR  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = TSR().to(dev)
amb = torch.from_numpy(face_ambiguous_cases()).to(dev)
# interior-ambiguous candidates: positive (or negative) corners containing a body diagonal pair that is not connected through
# positive (negative) edges -- Lewiner's cases 4, 6, 7, 10, 12, 13 and their complements
body = np.zeros(256, bool)
for case in range(256):
    for comp in (case, 255 - case):
        pos = {c for c in range(8) if (comp >> c) & 1}
        for c in list(pos):
            d = 7 - c
            if d in pos:
                # connected through edges inside pos?
                seen, todo = {c}, [c]
                while todo:
                    u = todo.pop()
                    for b in (1, 2, 4):
                        w = u ^ b
                        if w in pos and w not in seen:
                            seen.add(w)
                            todo.append(w)
                if d not in seen:
                    body[case] = True
body_t = torch.from_numpy(body).to(dev)
rows = []
for seed in (100, 101, 102, 103):
    tp = baked_triplane(seed).to(dev)
    dens = model.renderer.query_lattice(model.decoder, tp, R)
    thr = float(dens[:: max(1, R // 128)].median())
    cases = runtime.mc_cases(dens, sub=thr, sign=1.0).long().view(-1)
    active = (cases != 0) & (cases != 255)
    na = int(active.sum())
    rows.append({"seed": seed, "active_cells": na, "face_ambiguous": int((amb[cases] & active).sum()), "body_diagonal": int((body_t[cases] & active).sum()),
                 "either": int(((amb[cases] | body_t[cases]) & active).sum())})
tot = {k: sum(r[k] for r in rows) for k in ("active_cells", "face_ambiguous", "body_diagonal", "either")}
print(json.dumps({"resolution": R, "scenes": rows, "total": tot, "fraction_ambiguous": tot["either"] / max(1, tot["active_cells"])}))
