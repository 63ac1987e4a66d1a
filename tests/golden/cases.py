"""Golden vectors of the reference for this path: deterministic input builders shared by the generator
(tests/golden/make_golden.py, needs the compiled reference oracle/_ref) and by tests/test_golden.py (needs only the
committed tests/golden/golden_v1.npz).  The reference itself ships no known-answer vectors for these functions (checkasm is
differential), so these were produced by RUNNING the unmodified reference in the build container; the file
stores the reference's outputs plus a digest of every input so that a drift of the input generators is noticed.

Only numpy here: no library (oracle, reference, CUDA) is called from this module."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _libs import synth_luma, synth_sequence, worst_case_pair, PIXEL_W, PIXEL_H  # noqa: E402

GOLDEN = os.path.join(HERE, "golden_v1.npz")


def digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


# ---- x264_pixel_function_t: sad / ssd / satd [8 sizes], sa8d [16x16, 8x8] (common/pixel.c) ---------------------------
PIX_STRIDE, PIX_ROWS, PIX_NCAND = 64, 48, 48
PIX_COMBOS = [(m, ip) for m in (0, 1, 2) for ip in range(8)] + [(3, 0), (3, 3)]      # (metric, PIXEL_* index)


def pixel_case(pattern):
    rng = np.random.default_rng({"random": 11, "worst": 12}[pattern])
    n = PIX_STRIDE * PIX_ROWS
    if pattern == "random":
        a, b = rng.integers(0, 256, n, dtype=np.uint8), rng.integers(0, 256, n, dtype=np.uint8)
    else:
        a, b = worst_case_pair(n, rng)              # checkasm's overflow pattern, tools/checkasm.c:381-394
    fenc_off = (rng.integers(0, 16, PIX_NCAND) * PIX_STRIDE + rng.integers(0, 3, PIX_NCAND) * 16).astype(np.uint32)
    ref_off = (rng.integers(0, 16, PIX_NCAND) * PIX_STRIDE + rng.integers(0, 40, PIX_NCAND)).astype(np.uint32)
    return a, b, fenc_off, ref_off


# ---- x264_frame_init_lowres + border (common/mc.c:458-507, frame.c:627) and hpel_filter (mc.c:172-196) -----------------
LOWRES_WH = (100, 52)
HPEL_WH = (96, 80)


def lowres_case():
    w, h = LOWRES_WH
    return synth_luma(w, h, seed=w * h, kind="noise")


def hpel_case():
    w, h = HPEL_WH
    return synth_luma(w, h, seed=7 + w, kind="noise")


# ---- x264_me_search_ref (encoder/me.c:182-992) -----------------------------------------------------------------------
ME_W, ME_H, ME_MV_RANGE = 112, 96, 64
ME_GROUPS = [  # (method, subpel_refine, me_range, mbcmp_is_satd, weight)
    (0, 2, 16, 1, (0, 0, 0, 0)), (0, 5, 8, 1, (0, 0, 0, 0)), (1, 1, 16, 0, (0, 0, 0, 0)), (1, 4, 16, 1, (0, 0, 0, 0)),
    (1, 7, 16, 1, (1, 70, 6, -3)), (2, 3, 24, 1, (0, 0, 0, 0)), (2, 9, 32, 1, (0, 0, 0, 0)), (2, 6, 16, 1, (1, 55, 6, 4)),
    (3, 2, 8, 1, (0, 0, 0, 0)), (3, 7, 16, 1, (0, 0, 0, 0)),      # ESA (reference side: xref_me_search_frame, its own integral image)
    (3, 7, 16, 1, (1, 62, 6, 3)),                                  # ESA against a weighted reference: the ADS prefilter decides
    (4, 7, 16, 1, (0, 0, 0, 0)), (4, 2, 24, 1, (1, 75, 6, -2)), (4, 0, 8, 1, (0, 0, 0, 0)),      # TESA (fpelcmp = SATD)
]
ME_JOBS = 40


def me_content(seed):
    rng = np.random.default_rng(seed)
    ref_l = synth_luma(ME_W + 16, ME_H + 16, seed=1000 + seed)
    dx, dy = int(rng.integers(0, 9)), int(rng.integers(0, 9))
    fenc = ref_l[dy:dy + ME_H, dx:dx + ME_W].astype(np.int16) + rng.integers(-3, 4, (ME_H, ME_W))
    return np.clip(fenc, 0, 255).astype(np.uint8), np.ascontiguousarray(ref_l[4:4 + ME_H, 4:4 + ME_W])


def me_jobs(group_index):
    """plain-python job descriptions of one group (block, predictors, window, threshold)"""
    rng = np.random.default_rng(500 + group_index)
    jobs = []
    for _ in range(ME_JOBS):
        ip = int(rng.integers(0, 7))
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        bx = int(rng.integers(0, (ME_W - bw) // 4 + 1)) * 4
        by = int(rng.integers(0, (ME_H - bh) // 4 + 1)) * 4
        mvr = 4 * ME_MV_RANGE
        lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
        lim_max = [min(4 * (ME_W - bx - bw + 24), mvr - 1), min(4 * (ME_H - by - bh + 24), mvr - 1)]
        i_mvc = int(rng.integers(0, 5))
        spread = int(rng.choice([2, 12, 50]))
        mvp = rng.integers(-spread, spread + 1, 2)
        mvcs = rng.integers(-spread, spread + 1, (8, 2))
        if rng.random() < 0.3:
            mvp[:] = 0
        if rng.random() < 0.3 and i_mvc:
            mvcs[0] = mvp
        use_thresh = bool(rng.random() < 0.2)
        thresh = int(rng.integers(50, 3000))
        jobs.append(dict(ip=ip, bx=bx, by=by, lim_min=lim_min, lim_max=lim_max, i_mvc=i_mvc, mvp=[int(mvp[0]), int(mvp[1])],
                         mvc=mvcs.astype(np.int16), use_thresh=use_thresh, thresh=thresh))
    return jobs


# ---- x264_me_refine_bidir_satd (encoder/me.c:1027-1183) --------------------------------------------------------------------
BIDIR_CASES = [(1, 7), (0, 1)]          # (mbcmp is SATD, seed)
BIDIR_JOBS = 60


def bidir_case(ci):
    satd, seed = BIDIR_CASES[ci]
    rng = np.random.default_rng(seed)
    fenc_l, ref0_l = me_content(20 + seed)
    ref1_l = np.ascontiguousarray(np.roll(ref0_l, (3, -2), (0, 1)))
    jobs = []
    for _ in range(BIDIR_JOBS):
        ip = int(rng.integers(0, 7))
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        bx = int(rng.integers(0, (ME_W - bw) // 4 + 1)) * 4
        by = int(rng.integers(0, (ME_H - bh) // 4 + 1)) * 4
        mvr = 4 * ME_MV_RANGE
        lim_min = np.array([max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)])
        lim_max = np.array([min(4 * (ME_W - bx - bw + 24), mvr - 1), min(4 * (ME_H - by - bh + 24), mvr - 1)])
        spread = int(rng.choice([6, 30, 120]))
        mv0 = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max)
        mv1 = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max)
        mvp0, mvp1 = mv0 + rng.integers(-6, 7, 2), mv1 + rng.integers(-6, 7, 2)
        weight = int(rng.choice([32, 32, 21, 43, -10]))
        jobs.append(dict(ip=ip, bx=bx, by=by, lim_min=[int(v) for v in lim_min], lim_max=[int(v) for v in lim_max],
                         mv=[int(v) for v in np.concatenate([mv0, mv1])], mvp=[int(v) for v in np.concatenate([mvp0, mvp1])], weight=weight))
    return fenc_l, ref0_l, ref1_l, jobs


# ---- x264_me_refine_qpel / x264_me_refine_qpel_refdupe (encoder/me.c:800-814) ------------------------------------------------
REFINE_CASES = [(0, 1, 5, (0, 0, 0, 0)), (0, 0, 1, (0, 0, 0, 0)), (1, 1, 7, (1, 66, 6, -2)), (1, 1, 9, (0, 0, 0, 0))]   # (refdupe, satd, subme, weight)
REFINE_JOBS = 50


def refine_case(ci):
    rng = np.random.default_rng(70 + ci)
    fenc_l, ref_l = me_content(40 + ci)
    jobs = []
    for _ in range(REFINE_JOBS):
        ip = int(rng.integers(0, 7))
        bw, bh = PIXEL_W[ip], PIXEL_H[ip]
        bx = int(rng.integers(0, (ME_W - bw) // 4 + 1)) * 4
        by = int(rng.integers(0, (ME_H - bh) // 4 + 1)) * 4
        mvr = 4 * ME_MV_RANGE
        lim_min = [max(4 * (-bx - 24), -mvr), max(4 * (-by - 24), -mvr)]
        lim_max = [min(4 * (ME_W - bx - bw + 24), mvr - 1), min(4 * (ME_H - by - bh + 24), mvr - 1)]
        spread = int(rng.choice([4, 20, 90]))
        mv = np.clip(rng.integers(-spread, spread + 1, 2), lim_min, lim_max)
        mvp = mv + rng.integers(-6, 7, 2)
        use_thresh = bool(REFINE_CASES[ci][0] and rng.random() < 0.4)
        jobs.append(dict(ip=ip, bx=bx, by=by, lim_min=lim_min, lim_max=lim_max, mv=[int(mv[0]), int(mv[1])], mvp=[int(mvp[0]), int(mvp[1])],
                         cost=int(rng.integers(50, 6000)), ref_cost=int(rng.integers(0, 5)), use_thresh=use_thresh, thresh=int(rng.integers(50, 6000))))
    return fenc_l, ref_l, jobs


# ---- slicetype_frame_cost (encoder/slicetype.c:836-995) --------------------------------------------------------------
LA_CASES = [  # (preset, reference option string, (w, h)) ; the GPU / oracle parameters are stored in the fixture
    ("medium", "weightp=0:bframes=3", (112, 80)),
    ("medium", "weightp=0:bframes=3:subme=1:no-mbtree=1", (112, 80)),
    ("medium", "bframes=3", (112, 80)),                       # weight analysis on a fade
]
LA_NFR = 6
LA_REQUESTS = [(0, 0, 0), (0, 1, 1), (0, 2, 2), (0, 2, 1), (0, 3, 3), (0, 3, 1), (0, 3, 2), (1, 3, 2), (1, 2, 2),
               (2, 4, 3), (2, 4, 4), (1, 4, 4), (1, 4, 2), (1, 4, 3), (4, 4, 4), (3, 5, 4), (3, 5, 5), (2, 5, 5)]


def la_case(i, weighted):
    _, _, (w, h) = LA_CASES[i]
    frames = synth_sequence(w, h, LA_NFR, seed=w + h + i, cut_at=4)
    if weighted:
        frames = [np.clip(f.astype(np.float32) * (0.55 + 0.09 * k) + 3 * k, 0, 255).astype(np.uint8) for k, f in enumerate(frames)]
    rng = np.random.default_rng(1 + i)
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    qs = [rng.integers(180, 400, mbs).astype(np.uint16) for _ in range(LA_NFR)]
    return frames, qs


# ---- frame types of the reference ENCODER (x264_encoder_encode -> lookahead -> x264_slicetype_decide) ------------------
ST_CASES = [
    ("medium", "weightp=0:no-psy=1:bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 40, 17),
    ("medium", "weightp=0:no-psy=1:bframes=3:b-adapt=2:rc-lookahead=12:keyint=40", (96, 64), 36, 20),
    ("medium", "bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 40, 17),
]


def st_case(i):
    preset, opts, (w, h), n, cut = ST_CASES[i]
    frames = synth_sequence(w, h, n, seed=n + w + 3 * i, cut_at=cut)
    frames[cut + 7] = np.full_like(frames[0], 235)           # a two-frame flash, which must not become a scene cut
    frames[cut + 8] = np.full_like(frames[0], 235)
    for k in range(min(10, cut)):                             # a fade-in
        frames[k] = np.clip(frames[k].astype(np.float32) * (0.35 + 0.065 * k) + 2 * k, 0, 255).astype(np.uint8)
    return frames


# ---- x264_adaptive_quant_frame (encoder/ratecontrol.c:305-420) ----------------------------------------------------------
AQ_CASES = [((112, 80), 1, 1.0), ((100, 52), 2, 1.0), ((96, 64), 3, 1.3)]       # (size, aq-mode, aq-strength)


def aq_case(i):
    (w, h), _, _ = AQ_CASES[i]
    rng = np.random.default_rng(900 + i)
    luma = synth_luma(w, h, seed=900 + i)
    luma[: h // 3] = rng.integers(0, 256, (h // 3, w), dtype=np.uint8)
    luma[h // 3: h // 2, : w // 2] = 77
    cw, ch = (w + 1) // 2, (h + 1) // 2
    cb = rng.integers(100, 156, (ch, cw), dtype=np.uint8)
    cr = rng.integers(0, 256, (ch, cw), dtype=np.uint8)
    return np.ascontiguousarray(luma), cb, cr


# ---- MB-tree end to end: f_qp_offset of every non-B picture as the reference ENCODER used it (slicetype.c:1029-1184) -------
MBTREE_CASE = ("medium", "aq-mode=0:weightp=0:no-psy=1:bframes=3:rc-lookahead=10:keyint=30:min-keyint=3", (112, 80), 48, 29)


def mbtree_case():
    preset, opts, (w, h), n, cut = MBTREE_CASE
    return synth_sequence(w, h, n, seed=n + w + 1, cut_at=cut)
