"""Pins the CPU oracle (oracle/oracle.py) to golden vectors minted from the live reference.

Keypoints must be identical; float outputs are compared at 2e-6 abs (same ATen kernels, but the
oracle may batch differently from the reference module, which can change oneDNN blocking).
"""
import hashlib

import pytest
import torch

from oracle import oracle as O
from tests import golden_util as G
from tests import parity as PR


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("name", G.names("sparse"))
def test_sparse_matcher(name):
    g = G.load(name)
    kw = g["kwargs"]
    with torch.no_grad():
        k1, k2, p, d1, d2 = O.sparse_matcher(g["image1"], g["image2"], g["K"], return_descriptors=True, **kw)
        sc = O.shi_tomasi_score(g["image1"], kw.get("block_size", 3)).squeeze(1)
        mask = O.nms_mask(sc, kw.get("nms_radius", 3))
    assert _sha(sc) == g["score_sha"]
    assert _sha(mask) == g["mask_sha"]
    assert torch.equal(k1, g["kpts1"]) and torch.equal(k2, g["kpts2"])
    assert (d1 - g["desc1"]).abs().max() <= 2e-6
    assert (d2 - g["desc2"]).abs().max() <= 2e-6
    K = g["K"]
    assert (p - g["P"])[:, :K, :].abs().max() <= 2e-6
    assert ((p - g["P"])[:, K, K].abs() / g["P"][:, K, K]).max() <= 1e-5


@pytest.mark.parametrize("name", G.names("angle"))
def test_angle_matcher(name):
    g = G.load(name)
    kw = g["kwargs"]
    with torch.no_grad():
        k1, k2, p, d1, d2 = O.angle_matcher(g["image1"], g["image2"], g["K"], return_descriptors=True, **kw)
        a1 = O.angle_map(g["image1"])
        dk, ds, dd = O.angle_detector(g["image1"], g["K"], **kw)
    assert _sha(a1) == g["angle_sha"]
    assert torch.equal(k1, g["kpts1"]) and torch.equal(k2, g["kpts2"])
    assert (d1 - g["desc1"]).abs().max() <= 2e-6
    assert (d2 - g["desc2"]).abs().max() <= 2e-6
    K = g["K"]
    assert (p - g["P"])[:, :K, :].abs().max() <= 2e-6
    assert torch.equal(dk, g["det_kpts"])
    assert torch.equal(ds, g["det_scores"])
    assert (dd - g["det_desc"]).abs().max() <= 2e-6


@pytest.mark.parametrize("name", G.names("dense"))
def test_dense_matcher(name):
    g = G.load(name)
    kw = g["kwargs"]
    with torch.no_grad():
        k1, k2, p, d1, _ = O.dense_matcher(g["image1"], g["image2"], g["K"], return_descriptors=True, **kw)
        dkw = {k: v for k, v in kw.items() if k in ("block_size", "num_pairs", "binarize", "soft_binarize", "temperature")}
        sc, dmap = O.dense_detector(g["image1"], **dkw)
    assert _sha(sc) == g["score_sha"]
    assert _sha(dmap) == g["dmap_sha"]
    pb, pp, py, px = g["probe_idx"]
    assert torch.equal(dmap[pb, pp, py, px], g["probe_val"])
    assert torch.equal(k1, g["kpts1"]) and torch.equal(k2, g["kpts2"])
    assert (d1 - g["desc1"]).abs().max() <= 2e-6
    K = g["K"]
    assert (p - g["P"])[:, :K, :].abs().max() <= 2e-6


@pytest.mark.parametrize("name", G.names("sinkhorn"))
def test_sinkhorn(name):
    g = G.load(name)
    with torch.no_grad():
        p = O.sinkhorn(g["desc1"], g["desc2"], **g["kwargs"])
    assert torch.equal(p, g["P"])


@pytest.mark.parametrize("name", G.names("matches"))
def test_mutual_matches(name):
    """oracle.mutual_matches == the reference's MutualNearestNeighborMatcher (golden minted by make_golden_matches.py)."""
    g = G.load(name)
    src = G.load(g["source"])
    mk1, mk2, sc, valid = O.mutual_matches(src["P"], src["kpts1"], src["kpts2"], g["max_matches"], g["threshold"])
    assert torch.equal(sc, g["scores"]) and torch.equal(valid, g["valid"].bool())
    v = valid
    assert torch.equal(mk1[v], g["mk1"][v]) and torch.equal(mk2[v], g["mk2"][v])
    assert int(valid.sum()) > 0


@pytest.mark.parametrize("name", G.names("filters"))
def test_filter_rows(name):
    """oracle.filter_rows == the reference's SinkhornMatcherWithFilters on the golden P."""
    g = G.load(name)
    src = G.load(g["source"])
    pf, valid = O.filter_rows(src["P"], g["ratio_threshold"], g["dustbin_margin"])
    N = valid.shape[1]
    assert torch.equal(valid, g["valid"].bool()) and 0 < int(valid.sum()) < valid.numel()
    assert torch.equal(pf[:, :N, -1], g["dust_col"]) and torch.equal(pf.sum(dim=2), g["row_sums"])


def _essential_inputs(g):
    """(P, pts1_n, pts2_n, valid1, valid2) of one essential-matrix golden, batch dimension stripped."""
    K_inv = torch.linalg.inv(g["K"])
    if g["kind"] == "essential_grid":
        P = g["P"]
        shape = tuple(int(v) for v in g["image_shape"])
        return P, O.grid_points(P.shape[0] - 1, shape, K_inv), O.grid_points(P.shape[1] - 1, shape, K_inv), None, None
    k1, k2 = g["kpts1"][0], g["kpts2"][0]
    return g["P"][0], O.normalised_points(k1, K_inv), O.normalised_points(k2, K_inv), k1[:, 0] >= 0, k2[:, 0] >= 0


@pytest.mark.parametrize("name", G.names("essential_grid") + G.names("essential_module"))
def test_essential_matrix(name):
    """oracle.essential_matrix vs the reference's EssentialMatrixEstimator / ...WithEssentialMatrix head (goldens minted by
    make_golden_essential.py): the float32 restatement agrees to float rounding of the sums, the float64 one to the
    float32 error of the reference itself (5e-6 of max|E| measured; bound 5e-5)."""
    g = G.load(name)
    P, p1, p2, v1, v2 = _essential_inputs(g)
    kw = {k: v for k, v in g["kwargs"].items() if k in ("top_k", "n_iter", "n_iter_manifold")}
    E32 = O.essential_matrix(P, p1, p2, v1, v2, **kw)
    E64 = O.essential_matrix(P, p1, p2, v1, v2, dtype=torch.float64, **kw)
    scale = float(g["E"].abs().max())
    assert scale > 1.0
    assert float((E32 - g["E"]).abs().max()) <= 1e-5 * scale
    assert float((E64 - g["E"].double()).abs().max()) <= 5e-5 * scale
    # on the manifold: singular values (s, s, 0)
    sv = torch.linalg.svdvals(g["E"].double())
    assert float(sv[2]) <= 1e-4 * float(sv[0]) and abs(float(sv[0] - sv[1])) <= 1e-4 * float(sv[0])


def test_essential_matrix_no_weights():
    """every probability below the 0.01 threshold (the default epsilon = 1 at K = 512 does this): weights are all zero
    and the guarded divisions give E = 0 without NaN, as the reference does."""
    P = torch.full((33, 33), 0.005)
    pts = O.grid_points(32, (8, 8), torch.eye(3))
    E = O.essential_matrix(P, pts, pts)
    assert torch.equal(E, torch.zeros(3, 3))


def test_constant_image():
    g = G.load("sparse_constant_image")
    with torch.no_grad():
        k1, k2, p = O.sparse_matcher(g["image1"], g["image1"], g["K"])
    assert torch.equal(k1, g["kpts1"]) and (k1 == -1).all()
    assert torch.equal(p, g["P"])


def test_table_facts():
    for n in (256, 512):
        ox1, ox2, oy1, oy2, r, thr = O.bad_tables(n)
        assert ox1.shape == (n,) and thr.shape == (n,)
        assert int(r.min()) == 1 and int(r.max()) == 7
        for o in (ox1, ox2, oy1, oy2):
            assert o.min() >= -15 and o.max() <= 14
    with pytest.raises(ValueError):
        O.bad_tables(128)


# ------------------------------------------------------------------------------------------
# full-size goldens (tests/golden/make_golden_full.py): the sizes the benchmark runs
# ------------------------------------------------------------------------------------------
def test_dense_full_default():
    """BASELINE configs[1] at its real size: the oracle against the reference's own outputs."""
    g = G.load("dense_full_default")
    with torch.no_grad():
        k1, k2, p, d1, d2 = O.dense_matcher(g["image1"], g["image2"], 512, return_descriptors=True)
    assert torch.equal(k1, g["kpts1"]) and torch.equal(k2, g["kpts2"])
    assert (d1 - g["desc1"]).abs().max() <= 2e-6 and (d2 - g["desc2"]).abs().max() <= 2e-6
    assert (p - g["P"])[:, :512, :].abs().max() <= 2e-6
    m = G.load("matches_dense_full")
    mk1, mk2, sc, valid = O.mutual_matches(p, k1, k2, m["max_matches"], m["threshold"])
    assert torch.equal(valid, m["valid"].bool()) and int(valid.sum()) > 50
    assert (sc - m["scores"]).abs().max() <= 2e-6
    v = valid
    assert torch.equal(mk1[v], m["mk1"][v]) and torch.equal(mk2[v], m["mk2"][v])


def test_sparse_full_export():
    """The configuration the reference's export script ships, at full size (K = 1024: generic Sinkhorn path)."""
    g = G.load("sparse_full_export")
    with torch.no_grad():
        k1, k2, p, d1, d2 = O.sparse_matcher(g["image1"], g["image2"], 1024, return_descriptors=True, **g["kwargs"])
    assert torch.equal(k1, g["kpts1"]) and torch.equal(k2, g["kpts2"])
    assert (d1 - PR.full_descriptor_bits(g, 1)).abs().max() <= 2e-6 and (d2 - PR.full_descriptor_bits(g, 2)).abs().max() <= 2e-6
    PR.p_summary_ok(p, g, 2e-6)


def test_sparse_1080p_k2048():
    """BASELINE configs[4]: 1080x1920, K = 2048, both images, descriptors and P."""
    g = G.load("sparse_1080p_k2048")
    with torch.no_grad():
        k1, k2, p, d1, d2 = O.sparse_matcher(g["image1"], g["image2"], 2048, return_descriptors=True)
    assert torch.equal(k1, g["kpts1"]) and torch.equal(k2, g["kpts2"])
    rows = g["desc_rows"].long()
    assert (d1[0, rows] - g["desc1_sample"]).abs().max() <= 2e-6 and (d2[0, rows] - g["desc2_sample"]).abs().max() <= 2e-6
    assert (d1[0].double().sum(dim=-1).float() - g["desc1_rowsum"]).abs().max() <= 1e-4
    assert (d2[0].double().sum(dim=-1).float() - g["desc2_rowsum"]).abs().max() <= 1e-4
    PR.p_summary_ok(p, g, 2e-6)


def test_sinkhorn_with_scores():
    g = G.load("sinkhorn_with_scores")
    with torch.no_grad():
        p = O.sinkhorn(g["desc1"], g["desc2"], **g["kwargs"])
    N, M = g["desc1"].shape[1], g["desc2"].shape[1]
    assert torch.equal(p, g["P"])
    assert torch.equal(p[:, :N, :M].max(dim=-1).values, g["scores0"]) and torch.equal(p[:, :N, :M].max(dim=-2).values, g["scores1"])


@pytest.mark.parametrize("name", G.names("akaze"))
def test_akaze_maps_and_matcher(name):
    """the AKAZE matcher's detector maps (restated operator by operator) and everything behind them."""
    g = G.load(name)
    kw = g["kwargs"]
    with torch.no_grad():
        s1, o1 = O.akaze_maps(g["image1"])
        s2, o2 = O.akaze_maps(g["image2"])
        k1, k2, p, d1, d2 = O.maps_matcher(g["image1"], g["image2"], g["scores1"], g["scores2"], g["orient1"], g["orient2"],
                                           g["K"], return_descriptors=True, **kw)
    assert torch.equal(s1, g["scores1"]) and torch.equal(s2, g["scores2"])
    assert torch.equal(o1, g["orient1"]) and torch.equal(o2, g["orient2"])
    assert torch.equal(k1, g["kpts1"]) and torch.equal(k2, g["kpts2"])
    assert (d1 - g["desc1"]).abs().max() <= 2e-6 and (d2 - g["desc2"]).abs().max() <= 2e-6
    K = g["K"]
    assert (p - g["P"])[:, :K, :].abs().max() <= 2e-6


@pytest.mark.parametrize("name", G.names("ingest"))
def test_ingest_restatement_equals_opencv(name):
    """sample/visual_odometry.py:65-92 restated in integer numpy against what cv2 produced (golden minted with cv2)."""
    g = G.load(name)
    out = O.load_image_from_array(g["frame"].numpy(), g["height"], g["width"])
    ref = g["out"].numpy()
    assert out.shape == ref.shape and out.dtype == ref.dtype
    assert (out == ref).all()


def test_akaze_module_mirror_matches_oracle_on_cpu():
    """The product's AKAZE map module (torch operators, off the hot path) is the same arithmetic as the oracle restatement."""
    import onnx_image_processing_b200 as om
    g = G.load("akaze_small_default")
    with torch.no_grad():
        s, o = om.AKAZE()(g["image1"])
    assert torch.equal(s, g["scores1"]) and torch.equal(o, g["orient1"])


def test_config0_detector_480x640_k1000():
    """BASELINE configs[0] at its own size (one 480x640 image, max_keypoints=1000), both detector forms
    (tests/golden/make_golden_config0.py): the oracle against the live reference's outputs."""
    g = G.load("config0_detector_480x640_k1000")
    img, K = g["image1"], g["K"]
    with torch.no_grad():
        ak, asc, ad = O.angle_detector(img, K)
        sc, dmap = O.dense_detector(img)
        s3 = sc.squeeze(1)
        bk, bs = O.select_topk(s3, O.nms_mask(s3, g["nms_radius"]), K, g["threshold"], 0)
    assert torch.equal(ak, g["a_kpts"]) and torch.equal(asc, g["a_scores"])
    assert (ad - g["a_desc"]).abs().max() <= 2e-6
    assert torch.equal(sc, g["score_map"])
    assert torch.equal(bk, g["b_kpts"]) and torch.equal(bs, g["b_scores"])
    pp, py, px = g["probe_idx"].long()
    assert torch.equal(dmap[0, pp, py, px], g["probe_val"])
    yi, xi = bk[0, :, 0].long(), bk[0, :, 1].long()
    assert torch.equal(dmap[0][:, yi, xi].T, g["b_desc"])
