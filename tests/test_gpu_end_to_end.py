"""The reference's own model tests (tests/unit_tests.py:150-265), run through the drop-in API on the GPU engine:
CompositionalAnalysis(...) -> sample_hmc / sample_hmc_da / sample_nuts -> CAResult.summary_prepare()."""
import numpy as np
import pytest

from conftest import synthetic_dataset
from sccoda_b200 import datasets
from sccoda_b200.util import cell_composition_data as dat
from sccoda_b200.util import comp_ana as mod

pytestmark = pytest.mark.gpu


@pytest.fixture()
def data_salm():
    salm_df = datasets.haber().iloc[[0, 1, 2, 3, 8, 9], :]
    d = dat.from_pandas(salm_df, covariate_columns=["Mouse"])
    d.obs["Condition"] = d.obs["Mouse"].str.replace(r"_[0-9]", "", regex=True)
    return d


def _check_expected_samples(data, alpha_df, beta_df, cond="Condition[T.Salm]"):
    alphas_true = np.round(np.mean(data.X[:4], 0), 0)
    betas_true = np.round(np.mean(data.X[4:], 0), 0)
    final_alphas = np.round(alpha_df.loc[:, "Expected Sample"].tolist(), 0)
    final_betas = np.round(beta_df.loc[(cond,), "Expected Sample"].tolist(), 0)
    assert not any(np.abs(alphas_true - final_alphas) > 30), (alphas_true, final_alphas)
    assert not any(np.abs(betas_true - final_betas) > 30), (betas_true, final_betas)


def test_hmc(data_salm):
    np.random.seed(1234)
    model = mod.CompositionalAnalysis(data_salm, formula="Condition", reference_cell_type=5)
    result = model.sample_hmc(num_results=20000, num_burnin=5000, seed=5678)
    alpha_df, beta_df = result.summary_prepare()
    _check_expected_samples(data_salm, alpha_df, beta_df)
    # result contents of make_result (scCODA_model.py:596-638) and the double burn-in (15,000 retained draws)
    assert result.posterior["alpha"].shape == (1, 15000, 8) and result.posterior["beta"].shape == (1, 15000, 1, 8)
    assert set(result.sample_stats.data_vars) == {"target_log_prob", "diverging", "is_accepted", "step_size"}
    assert set(result.sampling_stats) == {"chain_length", "num_burnin", "acc_rate", "duration", "y_hat"}
    assert 0.4 < result.sampling_stats["acc_rate"] < 0.75               # tutorials: 50-61 %
    # the tutorial's call: Enterocyte is the one credible effect
    cred = result.credible_effects()
    assert bool(cred[("Condition[T.Salm]", "Enterocyte")])
    assert abs(beta_df.loc[("Condition[T.Salm]", "Enterocyte"), "Final Parameter"] - 1.35) < 0.15


def test_hmc_da(data_salm):
    np.random.seed(1234)
    model = mod.CompositionalAnalysis(data_salm, formula="Condition", reference_cell_type=5)
    result = model.sample_hmc_da(num_results=20000, num_burnin=5000, seed=5678)
    alpha_df, beta_df = result.summary_prepare()
    _check_expected_samples(data_salm, alpha_df, beta_df)
    assert set(result.sample_stats.data_vars) == {"target_log_prob", "diverging", "log_acc_ratio", "is_accepted",
                                                  "step_size"}


def test_nuts(data_salm):
    np.random.seed(1234)
    model = mod.CompositionalAnalysis(data_salm, formula="Condition", reference_cell_type=5)
    result = model.sample_nuts(num_results=2000, num_burnin=500, seed=5678)
    alpha_df, beta_df = result.summary_prepare()
    _check_expected_samples(data_salm, alpha_df, beta_df)
    assert set(result.sample_stats.data_vars) == {"target_log_prob", "leapfrogs_taken", "diverging", "energy",
                                                  "log_accept_ratio", "step_size", "reach_max_depth", "is_accepted"}


def test_multi_cond(data_salm):
    np.random.seed(1234)
    data_salm.obs["Condition2"] = np.random.randint(0, 2, len(data_salm.obs))
    model = mod.CompositionalAnalysis(data_salm, formula="Condition+Condition2", reference_cell_type=5)
    result = model.sample_hmc(num_results=20000, num_burnin=5000, seed=5678)
    alpha_df, beta_df = result.summary_prepare()
    _check_expected_samples(data_salm, alpha_df, beta_df)
    assert not any(beta_df.loc[("Condition2",), "Final Parameter"] != 0)      # no false positive on the random covariate


def test_multi_chain_fp64_and_target_log_prob(data_salm):
    np.random.seed(3)
    model = mod.CompositionalAnalysis(data_salm, formula="Condition", reference_cell_type="Goblet")
    # the reference's public log-density entry point (scCODA_model.py:756-760) at the SURVEY anchor state
    m5 = mod.CompositionalAnalysis(data_salm, formula="Condition", reference_cell_type=5)
    lp = m5.target_log_prob_fn(np.ones((1, 1)), np.zeros((1, 7)), np.zeros((1, 7)), np.zeros(8))
    assert abs(lp - (-260.45211233545143)) < 1e-8
    result = model.sample_hmc(num_results=3000, num_burnin=1000, verbose=False, num_chains=4, seed=1, dtype=np.float64)
    assert result.posterior["alpha"].shape == (4, 2000, 8)
    assert result.sample_stats["is_accepted"].shape == (4, 2000)
    assert result.effect_df.shape[0] == 8 and result.intercept_df.shape[0] == 8


def test_tutorial_intercepts_and_inclusion_probabilities_through_the_api(data_salm):
    """The reference-produced numbers of the tutorials (tests/test_oracle_statistical.py: tutorial_pins) through the
    drop-in API on the device: Haber Salmonella, reference Goblet, sample_hmc() defaults, 64 chains."""
    from test_oracle_statistical import tutorial_pins
    model = mod.CompositionalAnalysis(data_salm, formula="Condition", reference_cell_type="Goblet")
    result = model.sample_hmc(num_chains=64, seed=11, verbose=False)
    alpha = np.asarray(result.posterior["alpha"])                             # [64, 15000, 8]
    beta = np.asarray(result.posterior["beta"])                               # [64, 15000, 1, 8]
    assert alpha.shape == (64, 15000, 8) and beta.shape == (64, 15000, 1, 8)
    tutorial_pins(alpha.mean(1), (np.abs(beta[:, :, 0, :]) > 1e-3).mean(1))
    assert 0.45 < result.sampling_stats["acc_rate"] < 0.7


def test_sweep_inclusion_probabilities_and_credible_calls_match_oracle():
    """North-star check on sccoda.util.data_generation-style data: posterior inclusion probabilities and
    credible-effect calls of the device sweep (case/control kernel, fp32, 4 chains per dataset) against the fp64 CPU
    oracle run with the same sampler settings and independent random streams, within Monte-Carlo error."""
    from oracle import c_oracle, np_oracle as o
    from sccoda_b200 import scheduler as sch
    from sccoda_b200.util.data_generation import generate_sweep
    B, C, K, nr, nb = 12, 4, 20, 6000, 2000
    x, y, w_true = generate_sweep(B, K=K, n_per_group=10)
    cfg = sch.SweepConfig(method=0, num_results=nr, num_burnin=nb, chains=C, seed=11)
    res, shard = sch.run_sweep(x, y, K - 1, cfg)
    assert shard.out.kernel_kind == 3
    # oracle: same model, same settings, other streams; pooled over the chains exactly like shard_results
    ms = [o.prepare(x[b], y[b], K - 1) for b in range(B)]
    init = sch.initial_states(B, C, 1, K, seed=5)
    d, st_o = c_oracle.sample(np.stack([m.x for m in ms]), np.stack([m.y for m in ms]), np.stack([m.n_total for m in ms]),
                           [m.logp_const for m in ms], [m.ref for m in ms], init, 0, nr, nb, seed=99,
                           reject_conc_above=1e6)
    used = d[:, :, nb:, :]                                                    # second burn-in slice (:205-227)
    # Both sides sample the SAME target: the oracle runs with the concentration guard the fp32 kernels apply by default
    # (reject_conc_above = 1e6).  Without it some oracle chains are lost to the reference's precision-loss mode — the
    # fp64 lgamma differences of the reference graph vanish at astronomical concentrations and the density reads
    # +3e5 where its true value is -2e7 (tests/test_precision_loss.py) — and would have to be filtered after the fact.
    # No chain is dropped here.
    lp_o = st_o["target_log_prob"][:, :, nb:]                                 # [B,C,S]
    sane = lp_o.max(axis=2) < np.median(lp_o, axis=(1, 2))[:, None] + 500.0   # [B,C]
    assert sane.all(), sane.sum(1)
    sig, boff, iraw = used[..., 0:1], used[..., 1:K], used[..., K:2 * K - 1]
    with np.errstate(over="ignore"):
        beta = 1.0 / (1.0 + np.exp(-50.0 * iraw)) * sig * boff                # [B,C,S,K-1]; the reference column is last
    nzm = np.abs(beta) > 1e-3
    inc_chain = nzm.mean(2)                                                   # [B,C,K-1]
    wgt = sane[:, :, None] / sane.sum(1)[:, None, None]
    inc_o = np.concatenate([(inc_chain * wgt).sum(1), np.zeros((B, 1))], axis=1)          # [B,K]
    nz = (np.where(nzm, beta, 0.0).sum(2) * sane[:, :, None]).sum(1) / np.maximum((nzm.sum(2) * sane[:, :, None]).sum(1), 1)
    nz_o = np.concatenate([nz, np.zeros((B, 1))], axis=1)
    inc_d = res["inclusion_prob"]
    # inclusion probabilities: the indicator mixes slowly (+-0.05..0.1 per chain), so a single entry
    # pooled over 4 chains carries a Monte-Carlo error of ~0.05 on each side; the 12 x 19 entries together are sharp
    mean_c = (inc_chain * wgt).sum(1, keepdims=True)
    var_c = (((inc_chain - mean_c) ** 2) * sane[:, :, None]).sum(1) / np.maximum(sane.sum(1)[:, None] - 1, 1)
    se = np.sqrt(np.mean(var_c / sane.sum(1)[:, None]))                       # typical standard error of an entry
    diff = (inc_d - inc_o)[:, :K - 1]
    assert 0.01 < se < 0.12, se
    assert np.abs(diff).max() < 6 * np.sqrt(2) * se, (np.abs(diff).max(), se)
    assert np.sqrt(np.mean(diff ** 2)) < 2 * np.sqrt(2) * se, (np.sqrt(np.mean(diff ** 2)), se)
    assert abs(diff.mean()) < 0.02, diff.mean()
    assert np.corrcoef(inc_d[:, :K - 1].ravel(), inc_o[:, :K - 1].ravel())[0, 1] > 0.8
    # intercepts (alpha) and effects (beta): posterior means pooled over the chains, device against oracle, within the
    # Monte-Carlo error estimated from the spread of the oracle's chain means — a bias from the guard or from the fp32
    # arithmetic would show here first
    alpha_chain = used[..., -K:].mean(2)                                      # [B,C,K]
    se_a = float(np.sqrt(np.mean(alpha_chain.var(axis=1, ddof=1) / C)))
    d_alpha = res["mean"][:, -K:] - alpha_chain.mean(1)
    assert 1e-3 < se_a < 0.05, se_a
    assert np.abs(d_alpha).max() < 6 * np.sqrt(2) * se_a, (np.abs(d_alpha).max(), se_a)
    assert np.sqrt(np.mean(d_alpha ** 2)) < 2 * np.sqrt(2) * se_a, (np.sqrt(np.mean(d_alpha ** 2)), se_a)
    beta_chain = beta.mean(2)                                                 # [B,C,K-1]
    se_b = float(np.sqrt(np.mean(beta_chain.var(axis=1, ddof=1) / C)))
    d_beta = shard.problem.split_summary(shard.summary)["beta_mean"].mean(1).cpu().numpy()[:, :K - 1] - beta_chain.mean(1)
    assert np.abs(d_beta).max() < 6 * np.sqrt(2) * se_b, (np.abs(d_beta).max(), se_b)
    assert np.sqrt(np.mean(d_beta ** 2)) < 2 * np.sqrt(2) * se_b, (np.sqrt(np.mean(d_beta ** 2)), se_b)
    # credible-effect calls (result_classes.py:262-283): the same rule on both sides; they may differ only where the
    # inclusion probability sits at the threshold
    thr_o, final_o = sch.credible_effect_calls(inc_o, nz_o, 0.05)
    call_d, call_o = res["final_effect"] != 0, final_o != 0
    disagree = call_d != call_o
    assert disagree.mean() < 0.03, disagree.mean()
    for b, k in zip(*np.nonzero(disagree)):
        assert abs(inc_o[b, k] - thr_o[b]) < 0.15 or abs(inc_d[b, k] - res["threshold"][b]) < 0.15
    # the cell type that carries the true effect: same call on both sides, same size where it is called, and the
    # strong effects (log-fold 1.0) on reasonably abundant cell types are found
    found = 0
    for b in range(B):
        k = int(np.argmax(w_true[b]))
        assert call_d[b, k] == call_o[b, k] or abs(inc_o[b, k] - thr_o[b]) < 0.15
        if call_d[b, k] and call_o[b, k]:
            found += 1
            assert abs(res["final_effect"][b, k] - final_o[b, k]) < 0.15
            assert res["final_effect"][b, k] > 0
    assert found >= 3


def test_sweep_hdis_on_device_match_host_summary_logic():
    """SweepConfig.hdi_prob: intercept and effect HDIs of every dataset from the device-resident draws, against the
    numpy restatement of what CAResult.summary_prepare computes (result_classes.py:153-196) on the same draws."""
    from oracle import np_oracle as o
    from sccoda_b200 import scheduler as sch
    B, C, K, nr, nb = 3, 2, 6, 900, 300
    xs, ys = zip(*[synthetic_dataset(80 + b, N=12, K=K) for b in range(B)])
    refs = np.array([5, 0, 2])
    cfg = sch.SweepConfig(method=0, num_results=nr, num_burnin=nb, chains=C, seed=3, hdi_prob=0.94)
    res, shard = sch.run_sweep(np.stack(xs), np.stack(ys), refs, cfg)
    d = shard.out.draws.float().cpu().numpy()[:, :, nb:, :]                    # [B,C,S,P]
    a_hdi = res["alpha_hdi"].reshape(B, K, 2)
    e_hdi = res["effect_hdi"].reshape(B, K, 2)
    for b in range(B):
        pooled = d[b].reshape(-1, d.shape[-1])
        for k in range(K):
            assert np.allclose(a_hdi[b, k], o.hdi(pooled[:, -K + k].astype(np.float64), 0.94), atol=1e-6)
        sig, boff, iraw = pooled[:, 0:1], pooled[:, 1:K], pooled[:, K:2 * K - 1]
        ind = 1.0 / (1.0 + np.exp(-50.0 * iraw.astype(np.float64)))
        eff = np.where(ind >= 1e-3, sig * boff * ind, np.nan)
        j = 0
        for k in range(K):
            if k == refs[b]:
                assert np.all(e_hdi[b, k] == 0.0)
                continue
            lo, hi = o.hdi(eff[:, j], 0.94)
            assert np.allclose(e_hdi[b, k], [lo, hi], atol=2e-5, equal_nan=True), (b, k, e_hdi[b, k], lo, hi)
            j += 1


def test_reference_sweep_in_one_launch(data_salm):
    """BASELINE config 4 through the API: the 8 refits of the tutorial's reference sweep as one batch.  The tutorial
    (Modeling_options_and_result_analysis.ipynb:880-915) finds Enterocyte credible under 7 of 8 references (it is the
    reference itself in the eighth) and nothing else."""
    from sccoda_b200.util.comp_ana import reference_sweep
    times, res = reference_sweep(data_salm, "Condition", num_results=6000, num_burnin=2000, num_chains=4, seed=1)
    assert times.shape == (1, 8) and len(res["references"]) == 8
    row = times.iloc[0]
    assert row["Enterocyte"] >= 6
    assert (row.drop("Enterocyte") <= 1).all()
    assert res["inclusion_prob"].shape == (8, 8)
    for r in range(8):
        assert res["inclusion_prob"][r, r] == 0.0            # the reference column has no effect


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_predict_kernel_matches_numpy(dtype):
    """K5 (`sccoda_predict`): concentration / prediction per draw against the host formula of get_y_hat
    (scCODA_model.py:806-827) on a multi-covariate dataset, full and partial draw ranges."""
    import torch
    from conftest import random_states, synthetic_dataset
    from sccoda_b200.engine import DeviceProblem
    from sccoda_b200.model.scCODA_model import scCODAModel
    N, K, D, C, S, ref = 14, 7, 3, 2, 37, 2
    x, y = synthetic_dataset(5, N=N, K=K, D=D)
    prob = DeviceProblem.from_arrays(x, y, ref, chains=C, dtype=dtype)
    th = random_states(np.random.RandomState(1), D, K, C * S).reshape(1, C, S, -1)
    draws = torch.as_tensor(th.astype(dtype), device=prob.device)
    model = scCODAModel(reference_cell_type=ref, covariate_matrix=x, data_matrix=y, cell_types=[str(i) for i in range(K)],
                        covariate_names=[f"x{d}" for d in range(D)], formula="x")
    thd = draws.double().cpu().numpy()[0]
    K1 = K - 1
    states = [thd[..., :D, None], thd[..., D:D + D * K1].reshape(C, S, D, K1),
              thd[..., D + D * K1:D + 2 * D * K1].reshape(C, S, D, K1), thd[..., -K:]]
    model.get_y_hat(states, S, 0)
    conc_ref, pred_ref = states[-2], states[-1]
    tol = 2e-5 if dtype == np.float32 else 1e-12
    conc, pred = prob.predict(draws)
    assert conc.shape == (1, C, S, N, K) and conc.dtype == draws.dtype
    assert np.allclose(conc[0].double().cpu().numpy(), conc_ref, rtol=tol, atol=0)
    assert np.allclose(pred[0].double().cpu().numpy(), pred_ref, rtol=tol, atol=0)
    conc2, none = prob.predict(draws, 5, 19, prediction=False)
    assert none is None and torch.equal(conc2, conc[:, :, 5:19])
    none, pred2 = prob.predict(draws, S - 1, S, concentration=False)
    assert none is None and torch.equal(pred2, pred[:, :, S - 1:])
    with pytest.raises(RuntimeError):
        prob.predict(draws, 4, 4)


def test_result_predictive_arrays_are_generated_on_demand(tmp_path, data_salm):
    """`concentration` / `prediction` of a result are LazyArrays filled by the device on first use; values equal the
    host formula on the stored beta / alpha draws; a saved result holds plain arrays."""
    import pickle
    from sccoda_b200.util.result_classes import LazyArray
    model = mod.CompositionalAnalysis(data_salm, "Condition", reference_cell_type="Goblet")
    for chains in (1, 3):
        r = model.sample_hmc(num_results=600, num_burnin=200, verbose=False, seed=5, num_chains=chains)
        conc, pred = r.posterior["concentration"], r.posterior_predictive["prediction"]
        assert isinstance(conc, LazyArray) and isinstance(pred, LazyArray) and not conc.is_materialized
        assert conc.shape == (chains, 400, model.N, model.K) == pred.shape
        beta, alpha = np.asarray(r.posterior["beta"]), np.asarray(r.posterior["alpha"])
        ref = np.exp(np.einsum("nd,csdk->csnk", model.x, beta) + alpha[:, :, None, :])
        blk = conc.block(10, 20)
        assert blk.shape == (chains, 10, model.N, model.K) and not conc.is_materialized
        assert np.allclose(blk, ref[:, 10:20], rtol=2e-5)
        assert np.allclose(np.asarray(conc), ref, rtol=2e-5) and conc.is_materialized
        assert np.allclose(pred.sum(-1), model.n_total, rtol=1e-5)
        assert np.allclose(np.asarray(pred), model.n_total[:, None] * ref / ref.sum(-1, keepdims=True), rtol=3e-5)
        assert np.allclose(r.sampling_stats["y_hat"].sum(1), model.n_total)
    path = tmp_path / "r.pkl"
    r2 = model.sample_hmc(num_results=300, num_burnin=100, verbose=False, seed=6)
    r2.save(str(path))
    with open(path, "rb") as f:
        back = pickle.load(f)
    assert isinstance(back.posterior["concentration"], np.ndarray)
    assert back.posterior["concentration"].shape == (1, 200, model.N, model.K)
    assert np.allclose(back.posterior_predictive["prediction"].sum(-1), model.n_total, rtol=1e-5)
