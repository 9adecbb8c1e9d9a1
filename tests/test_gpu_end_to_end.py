"""The reference's own model tests (tests/unit_tests.py:150-265), run through the drop-in API on the GPU engine:
CompositionalAnalysis(...) -> sample_hmc / sample_hmc_da / sample_nuts -> CAResult.summary_prepare()."""
import numpy as np
import pytest

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
