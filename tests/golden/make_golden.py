#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz by RUNNING the unmodified reference (oracle/_ref/libx264ref.so, built from
/root/reference by oracle/Makefile.ref) on the deterministic cases of tests/golden/cases.py:

    make -f oracle/Makefile.ref -j8 && python tests/golden/make_golden.py

The reference cannot travel to machines without /root/reference's build products; these vectors can.  Stored per case: the
reference's outputs, the parameter structs it derived from its own option parsing, and a SHA-1 of every input."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from golden import cases as G, runners as R  # noqa: E402
import _libs  # noqa: E402


def main():
    assert _libs.have_ref(), "build oracle/_ref first (make -f oracle/Makefile.ref)"
    out = {}
    for pat in ("random", "worst"):
        out["pix_%s" % pat], out["pix_%s_in" % pat] = R.run_pixel("ref", pat)
    out["lowres"], out["lowres_in"] = R.run_lowres("ref")
    out["hpel"], out["hpel_in"] = R.run_hpel("ref")
    for gi in range(len(G.ME_GROUPS)):
        out["me_%d" % gi], out["me_%d_in" % gi] = R.run_me("ref", gi)
    for ci in range(len(G.REFINE_CASES)):
        out["refine_%d" % ci], out["refine_%d_in" % ci] = R.run_refine("ref", ci)
    for ci in range(len(G.BIDIR_CASES)):
        out["bidir_%d" % ci], out["bidir_%d_in" % ci] = R.run_bidir("ref", ci)
    for ci in range(len(G.LA_CASES)):
        res, pbytes, dig = R.run_la("ref", ci)
        for k, v in res.items():
            out["la_%d_%s" % (ci, k)] = v
        out["la_%d_params" % ci], out["la_%d_in" % ci] = pbytes, dig
    for ci in range(len(G.ST_CASES)):
        out["st_%d" % ci], out["st_%d_params" % ci], out["st_%d_in" % ci] = R.run_st("ref", ci)
    for ci in range(len(G.AQ_CASES)):
        (q, iq, st), dig = R.run_aq("ref", ci)
        out["aq_%d_qp" % ci], out["aq_%d_inv" % ci], out["aq_%d_stats" % ci], out["aq_%d_in" % ci] = q, iq, st, dig
    out["mbtree_types"], out["mbtree_qp"], out["mbtree_params"], out["mbtree_in"] = R.run_mbtree("ref")
    np.savez_compressed(G.GOLDEN, **out)
    print("wrote %s: %d arrays, %d bytes" % (G.GOLDEN, len(out), os.path.getsize(G.GOLDEN)))


if __name__ == "__main__":
    main()
