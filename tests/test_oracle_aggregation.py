"""CPU tests: the oracle restatement against (i) the reference's docstring known-answer vectors,
(ii) golden vectors produced by the reference's own files (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import aggregation as oa

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "aggregation_golden.json")))
KAT_J = torch.tensor([[-4.0, 1.0, 1.0], [6.0, 1.0, 1.0]])


def test_kat_mgda_docstrings():
    # mgda.py:57-86
    assert torch.allclose(oa.aggregate("mgda", KAT_J)[2], torch.tensor([0.0, 1.0, 1.0]), atol=1e-6)
    assert torch.allclose(oa.aggregate("mgda_ln", KAT_J)[2], torch.tensor([1.0, 1.0, 1.0]), atol=1e-6)
    ls = torch.tensor([0.5, 2.0])
    assert torch.allclose(oa.aggregate("mgda_gn", KAT_J, ls)[2], torch.tensor([3.4900, 1.0, 1.0]), atol=5e-5)
    assert torch.allclose(oa.aggregate("mgda_lgn", KAT_J, ls)[2], torch.tensor([4.1606, 1.0, 1.0]), atol=5e-5)


def test_kat_upgrad_docstring():
    # nupgrad.py:55-62 (torchjd UPGrad example) -- the ONLY reference-held pin for UPGrad
    G, w, g, _ = oa.aggregate("upgrad", KAT_J)
    assert torch.allclose(g, torch.tensor([0.2929, 1.9004, 1.9004]), atol=5e-5)
    assert torch.allclose(w, torch.tensor([1.110921, 0.789431]), atol=2e-6)


def test_kat_aligned_mtl():
    # torchjd AlignedMTL docstring + SURVEY 8c zero-row vector
    assert torch.allclose(oa.aggregate("aligned_mtl", KAT_J)[2], torch.tensor([0.2133, 0.9673, 0.9673]), atol=5e-5)
    Jz = torch.tensor([[-4.0, 1.0, 1.0], [0.0, 0.0, 0.0], [6.0, 1.0, 1.0]])
    _, w, g, info = oa.aggregate("aligned_mtl", Jz)
    assert torch.allclose(w, torch.tensor([0.3727, 0.0, 0.2722]), atol=5e-5)
    assert torch.allclose(g, torch.tensor([0.1422, 0.6449, 0.6449]), atol=5e-5)
    assert info["rank"] == 2
    for name in ("mgda_ln", "mgda_lgn"):
        _, w, g, _ = oa.aggregate(name, Jz, torch.tensor([0.5, 1.0, 2.0]))
        assert torch.equal(w, torch.tensor([0.0, 1.0, 0.0])) and float(g.abs().max()) == 0.0


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: c["tag"])
def test_golden_against_reference_files(case):
    G = torch.tensor(case["G"], dtype=torch.float32)
    losses = torch.tensor(case["losses"], dtype=torch.float32)
    for key, exp in case["out"].items():
        parts = key.split(":")
        if parts[0] in ("nupgrad", "pnupgrad"):
            continue                                   # covered by test_nupgrad_restatement_matches_reference_files
        if parts[0] == "aligned_mtl":
            w, _ = oa.aligned_mtl_weights(G, parts[1])
            # LAPACK ssyevd may differ in the last bits across CPUs -> tolerance, not bit equality
            np.testing.assert_allclose(w.numpy(), np.array(exp["w"], dtype=np.float32), rtol=2e-5, atol=1e-6)
        else:
            stable = len(parts) == 3
            w, count, gamma = oa.mgda_weights(G, parts[1], losses, stable=stable,
                                              min_eigenvalue_eps=1e-3 if stable else 1e-10)
            if stable:
                np.testing.assert_allclose(w.numpy(), np.array(exp["w"], dtype=np.float32), rtol=1e-4, atol=1e-5)
            else:
                # same float32 op sequence as the reference loop -> same iterates
                np.testing.assert_allclose(w.numpy(), np.array(exp["w"], dtype=np.float32), rtol=1e-6, atol=1e-7)
                assert count == exp["convergence_count"]
                assert gamma == pytest.approx(exp["gamma"], rel=1e-4, abs=1e-9)


@pytest.mark.parametrize("k", [2, 3, 4, 5, 8])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_upgrad_two_exact_solvers_agree_and_satisfy_kkt(k, seed):
    rng = np.random.default_rng(seed)
    J = rng.standard_normal((k, 50)) * np.logspace(0, -2, k)[:, None]
    if seed == 2:
        J[1] = 0.0
    G = torch.from_numpy(J @ J.T).float()
    w_gi = oa.upgrad_weights(G, solver="goldfarb_idnani")
    w_en = oa.upgrad_weights(G, solver="enumerate")
    np.testing.assert_allclose(w_gi.numpy(), w_en.numpy(), rtol=1e-6, atol=1e-7)
    H = oa.upgrad_prepare(G, 1e-4, 1e-4).double().numpy()
    for i in range(k):
        lo = np.zeros(k)
        lo[i] = 1.0 / k
        x = oa.qp_lower_bounds_goldfarb_idnani(H, lo)
        lam = H @ x
        assert np.all(x >= lo - 1e-10)                     # primal feasibility
        assert np.all(lam >= -1e-9 * np.abs(lam).max())    # dual feasibility
        assert abs(lam @ (x - lo)) <= 1e-9 * max(1.0, np.abs(lam).max())   # complementarity


@pytest.mark.parametrize("k", [2, 3, 5, 8])
def test_upgrad_qp_against_an_independent_library_solver(k):
    """Third, library-grade check of the UPGrad QP restatement (quadprog itself is not installable here): the dual-cone
    projection  min 1/2 v^T H v  s.t. v >= lo  is a bounded least-squares problem after a Cholesky factorisation
    (H = L L^T:  min 1/2 |L^T v|^2), which scipy.optimize.lsq_linear solves with an unrelated algorithm (BVLS)."""
    scipy_opt = pytest.importorskip("scipy.optimize")
    rng = np.random.default_rng(100 + k)
    J = rng.standard_normal((k, 40)) * np.logspace(0, -1.5, k)[:, None]
    G = torch.from_numpy(J @ J.T).float()
    H = oa.upgrad_prepare(G, 1e-4, 1e-4).double().numpy()
    Lc = np.linalg.cholesky(H)
    total = np.zeros(k)
    for i in range(k):
        lo = np.zeros(k)
        lo[i] = 1.0 / k
        res = scipy_opt.lsq_linear(Lc.T, np.zeros(k), bounds=(lo, np.full(k, np.inf)), method="bvls", tol=1e-14, max_iter=1000)
        assert res.success
        x = oa.qp_lower_bounds_goldfarb_idnani(H, lo)
        np.testing.assert_allclose(x, res.x, rtol=1e-6, atol=1e-9)
        total += res.x
    np.testing.assert_allclose(oa.upgrad_weights(G).numpy(), total.astype(np.float32), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("k", [2, 3, 5, 8])
def test_dualproj_restatement_two_solvers_kkt_and_library(k):
    """DualProj (torchjd, main.py:1221-1222): ONE QP with the whole mean-weight vector as lower bound.  Both exact solvers,
    the KKT conditions and scipy's BVLS agree; when no bound is active the result is u itself (u already lies in the dual cone)."""
    scipy_opt = pytest.importorskip("scipy.optimize")
    rng = np.random.default_rng(7 + k)
    J = rng.standard_normal((k, 30)) * np.logspace(0, -1, k)[:, None]
    J[0] = -0.7 * J[1] + 0.1 * J[0]                       # a conflicting pair: the projection has to move
    G = torch.from_numpy(J @ J.T).float()
    w_gi = oa.dualproj_weights(G, solver="goldfarb_idnani")
    w_en = oa.dualproj_weights(G, solver="enumerate")
    np.testing.assert_allclose(w_gi.numpy(), w_en.numpy(), rtol=1e-6, atol=1e-7)
    H = oa.upgrad_prepare(G, 1e-4, 1e-4).double().numpy()
    lo = np.full(k, 1.0 / k)
    x = w_gi.double().numpy()
    lam = H @ x
    assert np.all(x >= lo - 1e-6) and np.all(lam >= -1e-6 * np.abs(lam).max())
    res = scipy_opt.lsq_linear(np.linalg.cholesky(H).T, np.zeros(k), bounds=(lo, np.full(k, np.inf)), method="bvls", tol=1e-14)
    np.testing.assert_allclose(x, res.x, rtol=1e-5, atol=1e-7)
    w, _ = oa.weights_from_gramian("dualproj", G)
    assert torch.equal(w, w_gi)
    np.testing.assert_allclose(oa.dualproj_weights(torch.eye(k)).numpy(), lo, rtol=1e-6)      # orthogonal rows: u is feasible and optimal


def test_upgrad_zero_gramian_and_errors():
    w = oa.upgrad_weights(torch.zeros(3, 3))
    # trace < norm_eps -> G' = eps I -> projection of u_i e_i is itself -> w = 1/k each
    np.testing.assert_allclose(w.numpy(), np.full(3, 1 / 3), rtol=1e-6)
    with pytest.raises(ValueError):
        oa.mgda_weights(torch.eye(2), "bogus")
    with pytest.raises(RuntimeError):
        oa.mgda_weights(torch.eye(2), "loss", None)
    with pytest.raises(ValueError):
        oa.mgda_weights(torch.eye(2), "loss", torch.ones(3))
    with pytest.raises(ValueError):
        oa.aligned_mtl_weights(torch.eye(2), "bogus")
    with pytest.raises(ValueError):
        oa.aggregate("upgrad", torch.ones(3))


def test_similarity_from_gramian_matches_hook_formula():
    torch.manual_seed(0)
    J = torch.randn(4, 1000)
    w = torch.tensor([0.1, 0.5, 0.2, 0.9])
    ref = torch.nn.functional.cosine_similarity(J.T @ w, J.mean(0), dim=0).item()   # main.py:112-117
    got = oa.gradient_similarity_from_gramian(oa.gramian_fp64(J), w.tolist())
    assert got == pytest.approx(ref, abs=1e-5)


def test_nupgrad_restatement_matches_reference_files():
    """oracle.nupgrad_weights against utils/torchmoo/{nupgrad,pnupgrad}.py run behind the shim (golden)."""
    import json
    import os

    import numpy as np
    import torch

    from oracle import aggregation as oa

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "aggregation_golden.json")))
    for case in gold["cases"]:
        G = torch.tensor(case["G"], dtype=torch.float32)
        for key, mode in (("nupgrad", "min_l2"), ("pnupgrad:l2", "l2"), ("pnupgrad:min_l2", "min_l2")):
            w = oa.nupgrad_weights(G, mode=mode)
            np.testing.assert_allclose(w.numpy(), np.array(case["out"][key]["w"], dtype=np.float32), rtol=1e-6, atol=1e-7,
                                       err_msg=f"{case['tag']} {key}")
            w2 = oa.nupgrad_weights(G, mode=mode, solver="enumerate")        # independent exact solver
            np.testing.assert_allclose(w2.numpy(), w.numpy(), rtol=1e-6, atol=1e-7)


def test_comfort_beta_schedule_endpoints():
    from oracle import aggregation as oa

    assert oa.comfort_beta(1, 10) == 0.01 and oa.comfort_beta(10, 10) == 1.0 and oa.comfort_beta(1, 1) == 1.0
    assert 0.01 < oa.comfort_beta(5, 10) < 1.0
    assert oa.comfort_beta(5, 10, k=0) == 0.01 + 0.99 * (4 / 9)
