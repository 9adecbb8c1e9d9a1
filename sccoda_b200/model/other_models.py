"""
Comparison model `SimpleModel` — host-side mirror of `sccoda.model.other_models.SimpleModel`
(sccoda/model/other_models.py:31-282): a Dirichlet-multinomial model with Normal priors and no spike-and-slab,

    y | x   ~ DirichletMultinomial(concentration = exp(alpha + x beta), total = sum_k y)
    alpha_k ~ Normal(0, 5),   b_dk ~ Normal(0, 1) (k != reference),   beta_d,ref = 0.

It runs on the same device kernels as scCODAModel: with sigma_d = 1 and ind_raw = 1 (sigmoid(50) == 1 in fp32 and
fp64) the scCODA density IS this density plus a constant, so the sampler is launched with `freeze_spike_slab` —
sigma_d and ind_raw get zero momentum and zero gradient and never move — and the constant is taken off the traced
log-density.  The other comparison models of other_models.py (scdney, ALDEx2, ANCOM, ... : R and statsmodels
wrappers) are out of scope.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

from . import scCODA_model as dm
from ..util import result_classes as res


class SimpleModel(dm.CompositionalModel):
    """Simple Dirichlet-Multinomial model with normal priors (other_models.py:31-109)."""

    def __init__(self, reference_cell_type: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.reference_cell_type = reference_cell_type
        self.param_names = ["b", "alpha", "beta", "concentration", "prediction"]          # other_models.py:62
        D, K = self.D, self.K
        self.init_params = [np.random.normal(0, 1, (D, K - 1)), np.random.normal(0, 1, (K,))]   # :106-109

    # the engine's state vector [sigma_d | b_offset | ind_raw | alpha] for a SimpleModel state [b, alpha]
    def _flat_state(self, state) -> np.ndarray:
        b, alpha = (np.asarray(s, dtype=np.float64) for s in state)
        return np.concatenate([np.ones(self.D), b.reshape(-1), np.ones(b.size), alpha.reshape(-1)])

    def _split_draws(self, draws: np.ndarray) -> List[np.ndarray]:
        D, K = self.D, self.K
        lead = draws.shape[:-1]
        b = draws[..., D:D + D * (K - 1)].reshape(lead + (D, K - 1))
        return [b, draws[..., D + 2 * D * (K - 1):]]

    def init_states(self, num_chains: int) -> np.ndarray:
        D, K = self.D, self.K
        states = [self.init_params] + [[np.random.normal(0, 1, (D, K - 1)), np.random.normal(0, 1, (K,))]
                                       for _ in range(num_chains - 1)]
        return np.stack([self._flat_state(s) for s in states])

    @property
    def logp_offset(self) -> float:
        """Priors of the frozen parameters: HalfCauchy(0,1) at 1 and Normal(0,1) at 1 (scCODA_model.py:703-722)."""
        D, K = self.D, self.K
        return D * (math.log(2.0 / math.pi) - math.log(2.0)) + D * (K - 1) * (-0.5 - 0.5 * math.log(2.0 * math.pi))

    def target_log_prob_fn(self, b, alpha) -> float:
        """SimpleModel's joint log-density at one state (other_models.py:103-104), fp64 device kernel."""
        prob = self._device_problem(1, np.float64)
        lp, _ = prob.logp_grad(self._flat_state([b, alpha])[None, None])
        return float(lp.item()) - self.logp_offset

    def sample_hmc(self, num_results: int = int(20e3), num_burnin: int = int(5e3),
                   num_adapt_steps: Optional[int] = None, num_leapfrog_steps: Optional[int] = 10,
                   step_size: float = 0.01, verbose: bool = False, **engine_kwargs) -> res.CAResult:
        """HMC with simple step-size adaptation, target acceptance 0.8 (other_models.py:111-220).

        Traced statistics: target_log_prob, diverging, is_accepted, step_size."""
        if num_adapt_steps is None:
            num_adapt_steps = int(0.8 * num_burnin)                                   # :163-164
        kernel = dict(method="hmc", step_size=step_size, num_leapfrog_steps=num_leapfrog_steps,
                      num_adapt_steps=num_adapt_steps, target_accept_prob=0.8,        # :168
                      freeze_spike_slab=True, verbose=verbose, **engine_kwargs)
        num_chains = int(engine_kwargs.get("num_chains", 1))
        init = self.init_params if num_chains == 1 else self.init_states(num_chains)
        keys = ("target_log_prob", "diverging", "is_accepted", "step_size")

        def trace_fn(kr):
            out = {k: kr[k] for k in keys}
            out["target_log_prob"] = out["target_log_prob"] - self.logp_offset
            return out

        states, kernel_results, duration = self.sampling(num_results, num_burnin, kernel, init, trace_fn=trace_fn)
        states_burnin, sample_stats, acc_rate = self.get_chains_after_burnin(states, kernel_results, num_burnin)
        y_hat = self.get_y_hat(states_burnin, num_results, num_burnin, lazy=True)
        self._last_out = None
        sampling_stats = {"chain_length": num_results, "num_burnin": num_burnin, "acc_rate": acc_rate,
                          "duration": duration, "y_hat": y_hat}
        return self.make_result(states_burnin, sample_stats, sampling_stats)

    def get_y_hat(self, states_burnin: List[np.ndarray], num_results: int, num_burnin: int, lazy: bool = False
                  ) -> np.ndarray:
        """beta, concentration, prediction per draw are APPENDED to `states_burnin`; returns the posterior-mode count
        matrix (other_models.py:222-282), vectorised.  lazy=True: concentration / prediction are generated on the device
        on first use (see scCODAModel.get_y_hat)."""
        b, alphas = states_burnin[:2]
        beta_ = np.insert(b, self.reference_cell_type, 0.0, axis=-1)                  # :257-261
        lz = self._lazy_predictive(num_burnin, one_chain=alphas.ndim == 2) if lazy else None
        if lz is not None and lz[0].shape[:-2] == alphas.shape[:-1]:
            conc_, pred_ = lz
        else:
            conc_ = np.exp(np.einsum("nd,...dk->...nk", self.x, beta_) + alphas[..., None, :])      # :265-266
            pred_ = self.n_total[:, None] * conc_ / conc_.sum(axis=-1, keepdims=True)     # :268-271 DirMult mean
        states_burnin.extend([beta_, conc_, pred_])
        draw_axes = tuple(range(alphas.ndim - 1))
        concentration = np.exp(self.x @ beta_.mean(axis=draw_axes) + alphas.mean(axis=draw_axes))   # :279
        return concentration / concentration.sum(axis=1, keepdims=True) * self.n_total[:, None]

    def make_result(self, states_burnin, sample_stats: dict, sampling_stats: dict) -> res.CAResult:
        """other_models.py:186-220: posterior b / alpha / beta / concentration, predictive prediction."""
        params = dict(zip(self.param_names, states_burnin))
        ref = self.reference_cell_type
        cell_types_nb = self.cell_types[:ref] + self.cell_types[ref + 1:]
        one_chain = params["alpha"].ndim == 2
        def with_chain(a):
            if isinstance(a, res.LazyArray):
                return a.with_leading_axis() if one_chain else a
            return np.asarray(a)[None] if one_chain else np.asarray(a)
        dims = {"alpha": ["cell_type"], "b": ["covariate", "cell_type_nb"], "beta": ["covariate", "cell_type"],
                "concentration": ["sample", "cell_type"], "prediction": ["sample", "cell_type"]}
        coords = {"cell_type": res.Coord(self.cell_types), "cell_type_nb": res.Coord(cell_types_nb),
                  "covariate": res.Coord(self.covariate_names), "sample": res.Coord(range(self.y.shape[0]))}
        posterior = res.Dataset({k: with_chain(v) for k, v in params.items() if "prediction" not in k}, dims, coords)
        predictive = res.Dataset({"prediction": with_chain(params["prediction"])}, dims, coords)
        stats = res.Dataset({k: with_chain(v) for k, v in sample_stats.items()}, {}, coords)
        observed = res.Dataset({"y": self.y}, {"y": ["sample", "cell_type"]}, coords)
        model_specs = {"reference": self.reference_cell_type, "formula": self.formula}
        return res.CAResult(sampling_stats, model_specs, posterior=posterior, posterior_predictive=predictive,
                            sample_stats=stats, observed_data=observed)
