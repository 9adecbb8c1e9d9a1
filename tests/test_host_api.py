"""CPU tests of the host-side mirror of the reference API (no kernel launches)."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest

from sccoda_b200 import datasets
from sccoda_b200.util import cell_composition_data as dat
from sccoda_b200.util import comp_ana as mod
from sccoda_b200.util import data_generation as gen
from sccoda_b200.util import result_classes as res
from sccoda_b200.util.formula import design_matrix


# ---------------------------------------------------------------- reference golden vectors (tests/unit_tests.py:44-109)
def test_case_control_gen_golden_vector():
    np.random.seed(1234)
    data = gen.generate_case_control(1, 2, 1000, [2, 2], None, None, None)
    assert not any(np.abs(data.obs["x_0"] - [0, 0, 1, 1]) > 1e-5)
    assert np.array_equal(data.X, np.array([[74., 926.], [58., 942.], [32., 968.], [53., 947.]]))
    assert not any(np.abs(data.uns["b_true"] - np.array([-1.8508832, 0.7326526])) > 1e-5)
    assert np.array_equal(data.uns["w_true"], np.array([[0., 0.]]))


def test_change_functions_known_answers():
    np.random.seed(1234)
    b, w = gen.b_w_from_abs_change(np.array([600, 400]), 100, 1000)
    assert not any(np.abs(b - [-0.51082562, -0.91629073]) > 1e-5)
    assert not any(np.abs(w - [0.44183275, 0.]) > 1e-5)
    assert np.array_equal(gen.counts_from_first(600, 1000, 2), [600., 400.])


def test_from_pandas_shapes():
    """tests/unit_tests.py:114-128"""
    salm_df = datasets.haber().iloc[[0, 1, 2, 3, 8, 9], :]
    data_salm = dat.from_pandas(salm_df, covariate_columns=["Mouse"])
    data_salm.obs["Condition"] = data_salm.obs["Mouse"].str.replace(r"_[0-9]", "", regex=True)
    assert data_salm.X.shape == (6, 8) and data_salm.obs.shape == (6, 2)
    assert list(data_salm.var.index) == ["Endocrine", "Enterocyte", "Enterocyte.Progenitor", "Goblet", "Stem", "TA",
                                         "TA.Early", "Tuft"]


def test_generate_sweep_is_reproducible_and_shard_independent():
    x, y, w = gen.generate_sweep(6, K=7, n_per_group=3)
    x2, y2, w2 = gen.generate_sweep(2, K=7, n_per_group=3, dataset0=4)
    assert np.array_equal(y[4:], y2) and np.array_equal(w[4:], w2)
    assert y.shape == (6, 6, 7) and np.all(y.sum(2) == 5000) and np.all((w != 0).sum(1) == 1)


# ---------------------------------------------------------------- formula / factory
def _haber_salm():
    df = datasets.haber().iloc[[0, 1, 2, 3, 8, 9], :]
    d = dat.from_pandas(df, covariate_columns=["Mouse"])
    d.obs["Condition"] = d.obs["Mouse"].str.replace(r"_[0-9]", "", regex=True)
    return d


def test_design_matrix_treatment_coding():
    obs = pd.DataFrame({"Condition": ["Control", "Control", "Salm", "H.poly"], "dose": [0.5, 1.0, 2.0, 4.0],
                        "flag": [0, 1, 1, 0]})
    m, names = design_matrix("Condition", obs)
    assert names == ["Condition[T.H.poly]", "Condition[T.Salm]"]
    assert np.array_equal(m, [[0, 0], [0, 0], [0, 1], [1, 0]])
    m, names = design_matrix("C(Condition, Treatment('Salm'))", obs)
    assert names == ["C(Condition, Treatment('Salm'))[T.Control]", "C(Condition, Treatment('Salm'))[T.H.poly]"]
    assert np.array_equal(m, [[1, 0], [1, 0], [0, 0], [0, 1]])
    m, names = design_matrix("Condition + dose+flag", obs)
    assert names == ["Condition[T.H.poly]", "Condition[T.Salm]", "dose", "flag"]
    assert np.array_equal(m[:, 2], obs["dose"]) and np.array_equal(m[:, 3], obs["flag"])
    with pytest.raises(KeyError):
        design_matrix("nope", obs)


def test_design_matrix_interactions_transforms_and_no_intercept():
    """The formula language beyond sums of names, with patsy's conventions (column names, term order, the left-most
    factor of an interaction varying fastest, full-rank coding where the lower-order term is absent)."""
    obs = pd.DataFrame({"Condition": ["Control", "Control", "Salm", "H.poly", "Salm", "H.poly"],
                        "Sex": ["f", "m", "f", "m", "m", "f"], "dose": [0.5, 1.0, 2.0, 4.0, 8.0, 16.0],
                        "age": [1.0, 2.0, 3.0, 4.0, 5.0, 7.0]})
    cond = {lv: (obs["Condition"] == lv).to_numpy(float) for lv in ("Control", "H.poly", "Salm")}
    male = (obs["Sex"] == "m").to_numpy(float)
    # a*b = a + b + a:b; the interaction columns run through the levels of the LEFT factor first
    m, names = design_matrix("Condition * Sex", obs)
    assert names == ["Condition[T.H.poly]", "Condition[T.Salm]", "Sex[T.m]", "Condition[T.H.poly]:Sex[T.m]",
                     "Condition[T.Salm]:Sex[T.m]"]
    assert np.array_equal(m[:, 3], cond["H.poly"] * male) and np.array_equal(m[:, 4], cond["Salm"] * male)
    m2, names2 = design_matrix("Condition + Sex + Condition:Sex", obs)
    assert names2 == names and np.array_equal(m2, m)
    # categorical x numeric with both margins: treatment coding inside the interaction; numeric terms sort behind
    m, names = design_matrix("dose + Condition + Condition:dose", obs)
    assert names == ["Condition[T.H.poly]", "Condition[T.Salm]", "dose", "Condition[T.H.poly]:dose", "Condition[T.Salm]:dose"]
    assert np.array_equal(m[:, 4], cond["Salm"] * obs["dose"])
    # ... without the numeric margin every level gets its own slope (full-rank coding)
    m, names = design_matrix("Condition:dose", obs)
    assert names == ["Condition[Control]:dose", "Condition[H.poly]:dose", "Condition[Salm]:dose"]
    assert np.array_equal(m[:, 0], cond["Control"] * obs["dose"])
    # numeric interactions, powers and removal
    m, names = design_matrix("(dose + age)**2", obs)
    assert names == ["dose", "age", "dose:age"] and np.array_equal(m[:, 2], obs["dose"] * obs["age"])
    m, names = design_matrix("dose*age - dose:age", obs)
    assert names == ["dose", "age"]
    # transforms
    m, names = design_matrix("np.log(dose) + I(age**2 / 2) + center(age) + standardize(dose)", obs)
    assert names == ["np.log(dose)", "I(age**2 / 2)", "center(age)", "standardize(dose)"]
    assert np.allclose(m[:, 0], np.log(obs["dose"])) and np.allclose(m[:, 1], obs["age"] ** 2 / 2)
    assert np.allclose(m[:, 2], obs["age"] - obs["age"].mean())
    assert np.allclose(m[:, 3], (obs["dose"] - obs["dose"].mean()) / obs["dose"].std(ddof=0))
    # no intercept: patsy codes the first categorical factor with all its levels; the reference drops column 0 anyway
    for f in ("0 + Condition", "Condition - 1", "-1 + Condition"):
        m, names = design_matrix(f, obs)
        assert names == ["Condition[H.poly]", "Condition[Salm]"]
        assert np.array_equal(m, np.stack([cond["H.poly"], cond["Salm"]], axis=1))
    m, names = design_matrix("0 + Condition + Sex", obs)
    assert names == ["Condition[H.poly]", "Condition[Salm]", "Sex[T.m]"]
    # two categorical factors interacting without their margins: patsy's mixed coding is not reproduced
    with pytest.raises(NotImplementedError):
        design_matrix("Condition:Sex", obs)
    with pytest.raises(NotImplementedError):
        design_matrix("C(Condition, Sum)", obs)


def test_compositional_analysis_factory():
    data = _haber_salm()
    np.random.seed(0)
    model = mod.CompositionalAnalysis(data, formula="Condition", reference_cell_type=5)
    assert (model.N, model.D, model.K) == (6, 1, 8) and model.reference_cell_type == 5
    assert model.covariate_names == ["Condition[T.Salm]"]
    assert np.array_equal(model.x[:, 0], [0, 0, 0, 0, 1, 1])
    assert np.allclose(model.n_total, data.X.sum(1))
    # init state of the reference: sigma_d = 1, ind_raw = 0 (scCODA_model.py:763-768)
    assert np.all(model.init_params[0] == 1) and np.all(model.init_params[2] == 0)
    assert model.init_params[1].shape == (1, 7) and model.init_params[3].shape == (8,)
    assert mod.CompositionalAnalysis(data, "Condition", reference_cell_type="Goblet").reference_cell_type == 3
    with pytest.raises(NameError):
        mod.CompositionalAnalysis(data, "Condition", reference_cell_type="NoSuchType")
    with pytest.raises(NameError):
        mod.CompositionalAnalysis(data, "Condition", reference_cell_type=8)


def test_automatic_reference_and_pseudocount(capsys):
    data = _haber_salm()
    rel = data.X / data.X.sum(1, keepdims=True)
    expected = int(np.argmin(rel.var(0) / rel.mean(0)))
    model = mod.CompositionalAnalysis(data, "Condition")          # "automatic"
    out = capsys.readouterr().out
    assert model.reference_cell_type == expected and "Automatic reference selection!" in out
    d2 = _haber_salm()
    d2.X[0, 7] = 0
    m2 = mod.CompositionalAnalysis(d2, "Condition", reference_cell_type=5)
    assert "Zero counts encountered in data! Added a pseudocount of 0.5." in capsys.readouterr().out
    assert m2.y[0, 7] == 0.5 and m2.n_total[0] == d2.X[0].sum() + 0.5
    d3 = _haber_salm()
    d3.X[:, :] = 0
    d3.X[0, 0] = 1
    with pytest.raises(ValueError):
        mod.CompositionalAnalysis(d3, "Condition")


def test_input_dimension_check():
    from sccoda_b200.model.scCODA_model import scCODAModel
    with pytest.raises(ValueError):
        scCODAModel(reference_cell_type=0, covariate_matrix=np.zeros((3, 1)), data_matrix=np.ones((4, 2)),
                    cell_types=["a", "b"], covariate_names=["x"], formula="x")


# ---------------------------------------------------------------- result class on a hand-made posterior
def _fake_result(S=2000, seed=0):
    from sccoda_b200.model.scCODA_model import scCODAModel
    rs = np.random.RandomState(seed)
    np.random.seed(seed)
    K, D, N, ref = 4, 1, 6, 3
    y = rs.randint(20, 200, size=(N, K)).astype(float)
    x = np.array([0, 0, 0, 1, 1, 1.])[:, None]
    model = scCODAModel(reference_cell_type=ref, covariate_matrix=x, data_matrix=y, cell_types=list("ABCD"),
                        covariate_names=["cond"], formula="cond")
    sigma = np.abs(rs.randn(S, D, 1)) + 0.5
    boff = rs.randn(S, D, K - 1) * 0.1 + np.array([1.0, 0.0, 0.3])
    iraw = np.stack([np.full(S, 0.5), np.full(S, -0.5), np.where(rs.rand(S) < 0.4, 0.5, -0.5)], axis=1)[:, None, :]
    alpha = rs.randn(S, K) * 0.05 + np.array([1.0, 2.0, 0.5, 1.5])
    states = [sigma, boff, iraw, alpha]
    y_hat = model.get_y_hat(states, S, 0)
    stats = {"target_log_prob": rs.randn(S), "is_accepted": rs.rand(S) < 0.6}
    result = model.make_result(states, stats, {"chain_length": S, "num_burnin": 0, "acc_rate": 0.6, "duration": 1.0,
                                               "y_hat": y_hat})
    return model, result, states


def test_result_contents_and_summary_logic():
    model, result, states = _fake_result()
    S = 2000
    assert set(result.posterior.data_vars) == {"sigma_d", "b_offset", "ind_raw", "alpha", "ind", "b_raw", "beta",
                                               "concentration"}
    assert result.posterior["beta"].shape == (1, S, 1, 4) and result.posterior["concentration"].shape == (1, S, 6, 4)
    assert result.posterior_predictive["prediction"].shape == (1, S, 6, 4)
    assert np.all(result.posterior["beta"][..., 3] == 0)                       # reference column
    assert list(result.posterior.coords["cell_type_nb"].values) == ["A", "B", "C"]
    assert np.allclose(result.posterior_predictive["prediction"].sum(-1), model.n_total)
    eff, icpt = result.effect_df, result.intercept_df
    inc = eff["Inclusion probability"].to_numpy()
    assert inc[0] == 1.0 and inc[1] == 0.0 and abs(inc[2] - 0.4) < 0.05 and inc[3] == 0.0
    assert result.model_specs["threshold_prob"] == 1.0
    fp = eff["Final Parameter"].to_numpy()
    assert fp[0] != 0 and np.all(fp[1:] == 0)
    beta0 = result.posterior["beta"][0, :, 0, 0]
    assert abs(fp[0] - beta0.mean()) < 1e-9
    assert list(result.credible_effects()) == [True, False, False, False]
    # expected sample sums to the mean sample size; log2 fold change consistent
    y_bar = model.y.sum(1).mean()
    assert abs(icpt["Expected Sample"].sum() - y_bar) < 1e-6 and abs(eff["Expected Sample"].sum() - y_bar) < 1e-6
    assert np.allclose(eff["log2-fold change"], np.log2(eff["Expected Sample"].to_numpy() / icpt["Expected Sample"].to_numpy()))
    # az.summary-style statistics
    a = result.posterior["alpha"][0]
    assert np.allclose(icpt["Final Parameter"], np.round(a.mean(0), 3))
    assert np.allclose(icpt["SD"], np.round(a.std(0, ddof=1), 3))
    lo, hi = res.hdi(a[:, 0], 0.94)
    assert icpt["HDI 3%"].iloc[0] == round(lo, 3) and icpt["HDI 97%"].iloc[0] == round(hi, 3)
    # a laxer FDR admits the 40 % effect: threshold floor(0.4xx) ; set_fdr updates the stored frames
    result.set_fdr(0.5)
    assert result.effect_df["Final Parameter"].to_numpy()[2] != 0
    assert 0.35 < result.model_specs["threshold_prob"] < 0.45
    with pytest.raises(ValueError):
        result.credible_effects(est_fdr=1.5)


def test_result_printing_and_pickle(tmp_path, capsys):
    _, result, _ = _fake_result(S=300, seed=1)
    result.summary()
    result.summary_extended(hdi_prob=0.9)
    out = capsys.readouterr().out
    assert "Compositional Analysis summary:" in out and "Spike-and-slab threshold" in out and "HDI 5%" in out
    p = os.path.join(str(tmp_path), "res.pkl")
    result.save(p)
    with open(p, "rb") as f:
        back = pickle.load(f)
    assert np.array_equal(back.posterior["alpha"], result.posterior["alpha"])
    dist = result.distance_to_truth()
    assert dist.shape == (5, 5) and dist["Cell Type"].iloc[0] == "Total"


def test_inclusion_threshold_matches_oracle_restatement():
    from oracle import np_oracle as o
    rs = np.random.RandomState(3)
    for _ in range(20):
        S, D, K = 400, 2, 5
        beta = rs.randn(S, D, K) * (rs.rand(S, D, K) < rs.rand(1, D, K))
        inc, nz, thr, final = o.inclusion_and_effects(beta, 0.1)
        assert res.inclusion_threshold(inc, 0.1) == thr


def test_automatic_reference_batch_matches_per_dataset_rule():
    from sccoda_b200.util.comp_ana import automatic_reference, automatic_reference_batch
    from sccoda_b200.util.data_generation import generate_sweep
    _, y, _ = generate_sweep(24, K=9, n_per_group=6, n_total=400)
    y[3, :, 2] = 0                                   # a cell type that is absent everywhere cannot be the reference
    refs = automatic_reference_batch(y, 0.2)
    assert refs.shape == (24,) and refs[3] != 2
    assert [int(r) for r in refs] == [automatic_reference(y[b], 0.2) for b in range(24)]
    with pytest.raises(ValueError, match="large enough presence"):
        automatic_reference_batch(np.zeros((2, 4, 3)))


def test_lazy_array_quacks_like_an_array(tmp_path):
    """`LazyArray` (the on-demand posterior-predictive arrays of a result): nothing is fetched until somebody looks,
    ranges of draws can be fetched on their own, and the materialised value behaves like / pickles as an ndarray."""
    import pickle
    from sccoda_b200.util.result_classes import LazyArray
    full = np.arange(5 * 3 * 2, dtype=np.float64).reshape(5, 3, 2)
    calls = []

    def fetch(lo, hi):
        calls.append((lo, hi))
        return full[lo:hi]

    a = LazyArray(full.shape, np.float64, fetch)
    assert a.shape == (5, 3, 2) and a.ndim == 3 and a.size == 30 and len(a) == 5 and a.dtype == np.float64
    assert not a.is_materialized and calls == [] and "on demand" in repr(a)
    assert np.array_equal(a.block(1, 3), full[1:3]) and calls == [(1, 3)] and not a.is_materialized
    b = a.with_leading_axis()
    assert b.shape == (1, 5, 3, 2) and np.array_equal(b.block(2, 4), full[None, 2:4])
    assert np.array_equal(np.asarray(a), full) and a.is_materialized and calls[-1] == (0, 5)
    n_calls = len(calls)
    assert np.array_equal(a[1:, 0], full[1:, 0]) and a.sum() == full.sum() and np.array_equal(a.mean(axis=0), full.mean(axis=0))
    assert np.array_equal(a.block(0, 2), full[0:2]) and len(calls) == n_calls          # served from the materialised value
    assert np.array_equal(np.asarray(b), full[None])
    back = pickle.loads(pickle.dumps(LazyArray(full.shape, np.float64, fetch)))
    assert isinstance(back, np.ndarray) and np.array_equal(back, full)
    with pytest.raises(ValueError):
        np.asarray(LazyArray((6, 3, 2), np.float64, fetch))                            # fetched shape must match
