"""Golden vectors of the reference (tests/golden/golden_v1.npz, produced by tests/golden/make_golden.py from the unmodified
reference) against the CPU oracle (-m "not gpu") and against the CUDA path through the C ABI (-m gpu).  Unlike the
test_oracle_* files these need neither /root/reference nor oracle/_ref at run time.  Everything is integer: bit-exact."""
import numpy as np
import pytest

from golden import cases as G, runners as R

GOLD = np.load(G.GOLDEN)


def _check_inputs(name, dig):
    assert np.array_equal(GOLD[name + "_in"], dig), "%s: the input generator no longer reproduces the golden inputs" % name


@pytest.fixture(scope="module")
def ctx():
    import x264_b200 as x
    c = x.Context(0)
    yield c
    c.close()


BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def _ctx(request, backend):
    return request.getfixturevalue("ctx") if backend == "cuda" else None


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("pattern", ["random", "worst"])
def test_pixel_table(request, backend, pattern):
    got, dig = R.run_pixel(backend, pattern, _ctx(request, backend))
    _check_inputs("pix_" + pattern, dig)
    assert np.array_equal(got, GOLD["pix_" + pattern]), np.argwhere(got != GOLD["pix_" + pattern])[:5]


@pytest.mark.parametrize("backend", BACKENDS)
def test_lowres_planes(request, backend):
    got, dig = R.run_lowres(backend, _ctx(request, backend))
    _check_inputs("lowres", dig)
    assert np.array_equal(got, GOLD["lowres"])


@pytest.mark.parametrize("backend", BACKENDS)
def test_hpel_planes(request, backend):
    got, dig = R.run_hpel(backend, _ctx(request, backend))
    _check_inputs("hpel", dig)
    assert np.array_equal(got, GOLD["hpel"])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("gi", range(len(G.ME_GROUPS)))
def test_me_search(request, backend, gi):
    got, dig = R.run_me(backend, gi, _ctx(request, backend))
    _check_inputs("me_%d" % gi, dig)
    want = GOLD["me_%d" % gi]
    assert np.array_equal(got, want), (G.ME_GROUPS[gi], np.argwhere(got != want)[:5])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("ci", range(len(G.REFINE_CASES)))
def test_me_refine_qpel(request, backend, ci):
    got, dig = R.run_refine(backend, ci, _ctx(request, backend))
    _check_inputs("refine_%d" % ci, dig)
    want = GOLD["refine_%d" % ci]
    assert np.array_equal(got, want), (G.REFINE_CASES[ci], np.argwhere(got != want)[:5])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("ci", range(len(G.BIDIR_CASES)))
def test_me_refine_bidir(request, backend, ci):
    got, dig = R.run_bidir(backend, ci, _ctx(request, backend))
    _check_inputs("bidir_%d" % ci, dig)
    want = GOLD["bidir_%d" % ci]
    assert np.array_equal(got, want), (G.BIDIR_CASES[ci], np.argwhere(got != want)[:5])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("ci", range(len(G.LA_CASES)))
def test_lookahead_frame_cost(request, backend, ci):
    res, _, dig = R.run_la(backend, ci, GOLD["la_%d_params" % ci], _ctx(request, backend))
    _check_inputs("la_%d" % ci, dig)
    for k, v in res.items():
        want = GOLD["la_%d_%s" % (ci, k)]
        assert np.array_equal(v, want), (G.LA_CASES[ci], k, np.argwhere(v != want)[:5])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("ci", range(len(G.ST_CASES)))
def test_frame_types(request, backend, ci):
    got, _, dig = R.run_st(backend, ci, GOLD["st_%d_params" % ci], _ctx(request, backend))
    _check_inputs("st_%d" % ci, dig)
    want = GOLD["st_%d" % ci]
    assert np.array_equal(got, want), (G.ST_CASES[ci], [z for z in zip(got.tolist(), want.tolist()) if z[0] != z[1]][:6])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("ci", range(len(G.AQ_CASES)))
def test_adaptive_quant(request, backend, ci):
    (q, iq, st), dig = R.run_aq(backend, ci, _ctx(request, backend))
    _check_inputs("aq_%d" % ci, dig)
    assert np.array_equal(st, GOLD["aq_%d_stats" % ci]) and np.array_equal(iq, GOLD["aq_%d_inv" % ci])
    assert np.array_equal(q, GOLD["aq_%d_qp" % ci]), float(np.abs(q - GOLD["aq_%d_qp" % ci]).max())      # float, bit for bit


@pytest.mark.parametrize("backend", BACKENDS)
def test_mbtree_qp_offsets(request, backend):
    types, qp, _, dig = R.run_mbtree(backend, GOLD["mbtree_params"], _ctx(request, backend))
    _check_inputs("mbtree", dig)
    assert np.array_equal(types, GOLD["mbtree_types"])
    assert qp.shape == GOLD["mbtree_qp"].shape and np.array_equal(qp, GOLD["mbtree_qp"]), float(np.abs(qp - GOLD["mbtree_qp"]).max())
    assert np.abs(qp).max() > 0.5
