"""
Dirichlet-multinomial spike-and-slab model with a reference cell type — host-side mirror of
`sccoda.model.scCODA_model` (sccoda/model/scCODA_model.py:23-840).

Same class names, constructor arguments, sampler methods, defaults and result contents as the reference; the
TensorFlow-Probability machinery behind `sampling()` is replaced by the persistent sm_100a kernels of
`sccoda_b200.engine` (no TensorFlow, no CPU fallback).  New optional keyword arguments of the sampler methods:
`num_chains` (default 1, as the reference), `seed`, `device`, `dtype` (default float32 — the reference computes in
float64; `dtype=np.float64` selects the fp64 build of the kernels) and `reject_conc_above`: states with a
concentration exp(alpha_k + x.beta_k) above this limit are rejected (default 1e6 in fp32, 1e15 in fp64 — beyond
that the reference's own float64 differences lgamma(c+y) - lgamma(c) have no correct digit left and evaluate to the
bare multinomial constant, a spurious mode; `reject_conc_above=np.inf` with `dtype=np.float64` reproduces the
reference's arithmetic including that mode).

The model (scCODA_model.py:646-658):
    y | x      ~ DirichletMultinomial(concentration = exp(alpha + x beta), total = sum_k y)
    alpha_k    ~ Normal(0, 5)
    beta_dk    = sigmoid(50 ind_raw_dk) * sigma_d * b_offset_dk   (k != reference),   beta_d,ref = 0
    ind_raw_dk ~ Normal(0, 1),  b_offset_dk ~ Normal(0, 1),  sigma_d ~ HalfCauchy(0, 1)
"""
from __future__ import annotations

import time
from typing import List, Optional, Tuple

import numpy as np

from .. import _native as nat
from .._pack import apply_pseudocount, pack
from ..util import result_classes as res


class CompositionalModel:
    """Generic driver: holds the data, runs the samplers, packs results (scCODA_model.py:23-638)."""

    def __init__(self, covariate_matrix: np.ndarray, data_matrix: np.ndarray, cell_types: List[str],
                 covariate_names: List[str], formula: str, *args, **kwargs):
        self.x = np.array(covariate_matrix, dtype=np.float64)
        if self.x.ndim == 1:
            self.x = self.x[:, None]
        data_matrix = np.asarray(data_matrix)
        if np.count_nonzero(data_matrix) != np.size(data_matrix):               # scCODA_model.py:94-96
            print("Zero counts encountered in data! Added a pseudocount of 0.5.")
        self.y = apply_pseudocount(data_matrix)
        self.n_total = self.y.sum(axis=1)                                       # :99-100 (after the pseudocount)
        self.cell_types = cell_types
        self.covariate_names = covariate_names
        self.formula = formula
        self.N, self.D = self.x.shape
        self.K = self.y.shape[1]
        if self.N != self.y.shape[0]:                                           # :110-113
            raise ValueError("Wrong input dimensions X[{},:] != y[{},:]".format(self.x.shape[0], self.y.shape[0]))
        if self.N != len(self.n_total):
            raise ValueError("Wrong input dimensions X[{},:] != n_total[{}]".format(self.x.shape[0], len(self.n_total)))
        self._problems = {}

    # ------------------------------------------------------------------ device plumbing
    def _device_problem(self, chains: int, dtype, device=None):
        from ..engine import DeviceProblem
        key = (int(chains), np.dtype(dtype).name, device)
        if key not in self._problems:
            packed = pack(self.x, self.y, self.reference_cell_type, dtype=dtype, pseudocount=False)
            if packed.grouped and packed.Umax > 16 and chains < 1184:
                # a handful of chains and more than 16 covariate groups (a continuous covariate): the planner keeps such
                # runs on the generic kernels with many warps per chain (capi.cu: make_plan), and there the element
                # layout is the faster one (17.9 ms against 23.7 ms per 500 HMC steps at 32 groups)
                packed = pack(self.x, self.y, self.reference_cell_type, dtype=dtype, pseudocount=False, grouped=False)
            self._problems[key] = DeviceProblem(packed, chains=chains, device=device)
        return self._problems[key]

    def _flat_state(self, state) -> np.ndarray:
        return np.concatenate([np.asarray(s, dtype=np.float64).reshape(-1) for s in state])

    def _split_draws(self, draws: np.ndarray) -> List[np.ndarray]:
        """[..., S, P] -> [sigma_d [..., S, D, 1], b_offset [..., S, D, K-1], ind_raw [..., S, D, K-1], alpha [..., S, K]]"""
        D, K = self.D, self.K
        lead = draws.shape[:-1]
        o = 0
        sigma = draws[..., o:o + D].reshape(lead + (D, 1)); o += D
        boff = draws[..., o:o + D * (K - 1)].reshape(lead + (D, K - 1)); o += D * (K - 1)
        iraw = draws[..., o:o + D * (K - 1)].reshape(lead + (D, K - 1)); o += D * (K - 1)
        alpha = draws[..., o:o + K]
        return [sigma, boff, iraw, alpha]

    # ------------------------------------------------------------------ sampling
    def sampling(self, num_results: int, num_burnin: int, kernel: dict, init_state, trace_fn=None
                 ) -> Tuple[List[np.ndarray], dict, float]:
        """Run `num_burnin + num_results` transitions and return the last `num_results` states and their kernel
        statistics (what tfp.mcmc.sample_chain does for the reference, scCODA_model.py:115-169).

        kernel
            dict describing the transition kernel: method ("hmc" | "hmc_da" | "nuts"), step_size, num_leapfrog_steps,
            max_tree_depth, num_adapt_steps, num_chains, seed, dtype, device, verbose, reject_conc_above.
        init_state
            list of the four state parts (one chain) or an array [num_chains, P].
        trace_fn
            optional callable(stats_dict) -> dict selecting / renaming the traced statistics.
        Returns states (list of 4 arrays, [num_results, ...] for one chain, [num_chains, num_results, ...]
        otherwise), kernel_results (dict of arrays) and the sampling duration in seconds.
        """
        method = {"hmc": nat.METHOD_HMC, "hmc_da": nat.METHOD_HMC_DA, "nuts": nat.METHOD_NUTS}[kernel["method"]]
        chains = int(kernel.get("num_chains", 1))
        dtype = np.dtype(kernel.get("dtype", np.float32))
        seed = kernel.get("seed")
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        prob = self._device_problem(chains, dtype, kernel.get("device"))
        if isinstance(init_state, (list, tuple)):
            init = self._flat_state(init_state)[None, :]
        else:
            init = np.asarray(init_state, dtype=np.float64).reshape(-1, prob.P)
        if init.shape[0] != chains:
            if init.shape[0] != 1:
                raise ValueError(f"init_state has {init.shape[0]} rows but num_chains={chains}")
            init = np.broadcast_to(init, (chains, prob.P))
        smp = prob.make_sampler(method, int(num_results), int(num_burnin), kernel.get("num_adapt_steps"),
                                int(kernel.get("num_leapfrog_steps", 10)), int(kernel.get("max_tree_depth", 10)),
                                float(kernel.get("step_size", 0.01)), float(kernel.get("target_accept_prob", 0.75)),
                                seed=int(seed), freeze_spike_slab=bool(kernel.get("freeze_spike_slab", False)),
                                reject_conc_above=-1.0 if np.isinf(kernel.get("reject_conc_above", 0.0))
                                else float(kernel.get("reject_conc_above", 0.0)))
        total = num_burnin + num_results
        verbose = bool(kernel.get("verbose", False))
        progress, chunk = None, None
        if verbose:
            chunk = max(1, total // 20)

            def progress(done, tot):
                print("\r{:3d}%|{:<20s}| {}/{}".format(100 * done // tot, "#" * (20 * done // tot), done, tot),
                      end="" if done < tot else "\n", flush=True)

        start = time.time()
        out = prob.sample(init[None].astype(dtype), smp, chunk=chunk, progress=progress)
        draws = out.draws[0].double().cpu().numpy()                  # [C, S, P]
        stats = out.stats[0].double().cpu().numpy()                  # [C, S, NSTAT]
        duration = time.time() - start
        print("MCMC sampling finished. ({:.3f} sec)".format(duration))

        kr = {name: stats[..., i] for i, name in enumerate(nat.STAT_NAMES)}
        for name in ("is_accepted", "diverging", "reach_max_depth"):
            kr[name] = kr[name] > 0.5
        kr["leapfrogs_taken"] = kr["leapfrogs_taken"].astype(np.int64)
        if trace_fn is not None:
            kr = trace_fn(kr)
        states = self._split_draws(draws)
        if chains == 1:
            states = [s[0] for s in states]
            kr = {k: v[0] for k, v in kr.items()}
        self._last_out = (prob, out)                                  # device-resident draws, for the lazy predictive
        self._last_run = {"kernel_ms": out.kernel_ms, "warps_per_chain": out.warps_per_chain, "grid": out.grid,
                          "chains": chains, "launches": out.launches}
        return states, kr, duration

    def get_chains_after_burnin(self, samples: List[np.ndarray], kernel_results: dict, num_burnin: int,
                                is_nuts: bool = False) -> Tuple[List[np.ndarray], dict, float]:
        """Cut a further `num_burnin` draws from the traced chain and compute the acceptance rate over the whole
        traced chain (scCODA_model.py:171-227 — the reference's double burn-in is kept on purpose)."""
        draw_axis = samples[-1].ndim - 2          # alpha is [S, K] (one chain) or [C, S, K]
        cut = (slice(None),) * draw_axis + (slice(num_burnin, None),)
        states_burnin = [np.asarray(s)[cut] for s in samples]
        stats = {k: np.asarray(v)[cut] for k, v in kernel_results.items()}
        if is_nuts:
            p_accept = float(np.mean(np.exp(kernel_results["log_accept_ratio"])))
        else:
            acc = np.asarray(kernel_results["is_accepted"])
            p_accept = float(acc.sum() / acc.size)
        print('Acceptance rate: %0.1f%%' % (100 * p_accept))
        return states_burnin, stats, p_accept

    def _run(self, method: str, num_results, num_burnin, num_adapt_steps, step_size, verbose, trace_keys,
             num_leapfrog_steps=10, max_tree_depth=10, num_chains=1, seed=None, dtype=np.float32, device=None,
             reject_conc_above=0.0):
        if num_adapt_steps is None:
            num_adapt_steps = int(0.8 * num_burnin)                              # scCODA_model.py:284-285
        kernel = dict(method=method, step_size=step_size, num_leapfrog_steps=num_leapfrog_steps,
                      max_tree_depth=max_tree_depth, num_adapt_steps=num_adapt_steps, target_accept_prob=0.75,
                      num_chains=num_chains, seed=seed, dtype=dtype, device=device, verbose=verbose,
                      reject_conc_above=reject_conc_above)
        init = self.init_params if num_chains == 1 else self.init_states(num_chains)
        states, kernel_results, duration = self.sampling(
            num_results, num_burnin, kernel, init, trace_fn=lambda kr: {new: kr[old] for new, old in trace_keys})
        states_burnin, sample_stats, acc_rate = self.get_chains_after_burnin(states, kernel_results, num_burnin,
                                                                             is_nuts=(method == "nuts"))
        y_hat = self.get_y_hat(states_burnin, num_results, num_burnin, lazy=True)
        self._last_out = None                                         # the lazy arrays hold what they need
        sampling_stats = {"chain_length": num_results, "num_burnin": num_burnin, "acc_rate": acc_rate,
                          "duration": duration, "y_hat": y_hat}
        return self.make_result(states_burnin, sample_stats, sampling_stats)

    def _lazy_predictive(self, num_burnin: int, one_chain: bool):
        """`concentration` and `prediction` per draw (after the second burn-in) as arrays that are generated on the
        device on first use from the draws of the last `sampling` call (K5, `DeviceProblem.predict`); None when there
        is no device-resident run to refer to."""
        last = getattr(self, "_last_out", None)
        if last is None:
            return None
        prob, out = last
        C, S = int(out.draws.shape[1]), int(out.draws.shape[2])
        n_draws = S - int(num_burnin)
        if n_draws <= 0:
            return None
        N, K = prob.N, prob.K
        per_draw = C * N * K * out.draws.element_size()
        step = max(1, int((1 << 29) // max(per_draw, 1)))             # device buffers of at most ~512 MB per call

        def fetcher(which):
            def fetch(lo, hi):
                res_ = np.empty((C, hi - lo, N, K), dtype=np.float64)
                for s0 in range(lo, hi, step):
                    s1 = min(hi, s0 + step)
                    conc, pred = prob.predict(out.draws, num_burnin + s0, num_burnin + s1,
                                              concentration=which == "concentration", prediction=which == "prediction")
                    t = conc if which == "concentration" else pred
                    res_[:, s0 - lo:s1 - lo] = t[0].double().cpu().numpy()
                return res_[0] if one_chain else res_
            return fetch

        shape = (n_draws, N, K) if one_chain else (C, n_draws, N, K)
        axis = 0 if one_chain else 1
        return (res.LazyArray(shape, np.float64, fetcher("concentration"), axis),
                res.LazyArray(shape, np.float64, fetcher("prediction"), axis))

    def sample_hmc(self, num_results: int = int(20e3), num_burnin: int = int(5e3),
                   num_adapt_steps: Optional[int] = None, num_leapfrog_steps: Optional[int] = 10,
                   step_size: float = 0.01, verbose: bool = True, **engine_kwargs) -> res.CAResult:
        """HMC with simple step-size adaptation (target 0.75) — scCODA_model.py:229-332.

        Traced statistics: target_log_prob, diverging (log_accept_ratio < -1000), is_accepted, step_size."""
        keys = [("target_log_prob", "target_log_prob"), ("diverging", "diverging"), ("is_accepted", "is_accepted"),
                ("step_size", "step_size")]
        return self._run("hmc", num_results, num_burnin, num_adapt_steps, step_size, verbose, keys,
                         num_leapfrog_steps=num_leapfrog_steps, **engine_kwargs)

    def sample_hmc_da(self, num_results: int = int(20e3), num_burnin: int = int(5e3),
                      num_adapt_steps: Optional[int] = None, num_leapfrog_steps: Optional[int] = 10,
                      step_size: float = 0.01, verbose: bool = True, **engine_kwargs) -> res.CAResult:
        """HMC with dual-averaging step-size adaptation (Nesterov 2009) — scCODA_model.py:334-438.

        Traced statistics: target_log_prob, diverging, log_acc_ratio, is_accepted, step_size (the averaged
        iterate exp(log_averaging_step), as in the reference)."""
        keys = [("target_log_prob", "target_log_prob"), ("diverging", "diverging"),
                ("log_acc_ratio", "log_accept_ratio"), ("is_accepted", "is_accepted"), ("step_size", "step_size")]
        return self._run("hmc_da", num_results, num_burnin, num_adapt_steps, step_size, verbose, keys,
                         num_leapfrog_steps=num_leapfrog_steps, **engine_kwargs)

    def sample_nuts(self, num_results: int = int(10e3), num_burnin: int = int(5e3),
                    num_adapt_steps: Optional[int] = None, max_tree_depth: int = 10, step_size: float = 0.01,
                    verbose: bool = True, **engine_kwargs) -> res.CAResult:
        """No-U-turn sampling with dual-averaging adaptation — scCODA_model.py:440-567.

        Traced statistics: target_log_prob, leapfrogs_taken, diverging, energy, log_accept_ratio, step_size,
        reach_max_depth, is_accepted."""
        keys = [(k, k) for k in ("target_log_prob", "leapfrogs_taken", "diverging", "energy", "log_accept_ratio",
                                 "step_size", "reach_max_depth", "is_accepted")]
        return self._run("nuts", num_results, num_burnin, num_adapt_steps, step_size, verbose, keys,
                         max_tree_depth=max_tree_depth, **engine_kwargs)

    # ------------------------------------------------------------------ result packing
    def make_result(self, states_burnin, sample_stats: dict, sampling_stats: dict) -> res.CAResult:
        """Chain states -> CAResult with the groups / dims / coords of the reference (scCODA_model.py:569-638)."""
        params = dict(zip(self.param_names, states_burnin))
        ref = self.reference_cell_type
        cell_types_nb = (self.cell_types[:ref] + self.cell_types[ref + 1:]) if ref is not None else self.cell_types
        one_chain = params["alpha"].ndim == 2

        def with_chain(a):
            if isinstance(a, res.LazyArray):
                return a.with_leading_axis() if one_chain else a
            a = np.asarray(a)
            return a[None] if one_chain else a

        dims = {"alpha": ["cell_type"], "sigma_d": ["covariate"], "b_offset": ["covariate", "cell_type_nb"],
                "ind_raw": ["covariate", "cell_type_nb"], "ind": ["covariate", "cell_type_nb"],
                "b_raw": ["covariate", "cell_type_nb"], "beta": ["covariate", "cell_type"],
                "concentration": ["sample", "cell_type"], "prediction": ["sample", "cell_type"]}
        coords = {"cell_type": res.Coord(self.cell_types), "cell_type_nb": res.Coord(cell_types_nb),
                  "covariate": res.Coord(self.covariate_names), "sample": res.Coord(range(self.y.shape[0]))}
        posterior = res.Dataset({k: with_chain(v) for k, v in params.items() if "prediction" not in k}, dims, coords)
        predictive = res.Dataset({"prediction": with_chain(params["prediction"])} if "prediction" in params else {},
                                 dims, coords)
        stats = res.Dataset({k: with_chain(v) for k, v in sample_stats.items()}, {}, coords)
        observed = res.Dataset({"y": self.y}, {"y": ["sample", "cell_type"]}, coords)
        model_specs = {"reference": self.reference_cell_type, "formula": self.formula}
        return res.CAResult(sampling_stats, model_specs, posterior=posterior, posterior_predictive=predictive,
                            sample_stats=stats, observed_data=observed)


class scCODAModel(CompositionalModel):
    """The standard scCODA model with a fixed reference cell type (scCODA_model.py:641-840)."""

    def __init__(self, reference_cell_type: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.reference_cell_type = reference_cell_type
        self.param_names = ["sigma_d", "b_offset", "ind_raw", "alpha", "ind", "b_raw", "beta", "concentration",
                            "prediction"]                                        # scCODA_model.py:692-693
        # MCMC starting values (scCODA_model.py:763-768): sigma_d = 1, b_offset ~ N(0,1), ind_raw = 0, alpha ~ N(0,1)
        D, K = self.D, self.K
        self.init_params = [np.ones((D, 1)), np.random.normal(0, 1, (D, K - 1)), np.zeros((D, K - 1)),
                            np.random.normal(0, 1, (K,))]

    @property
    def P(self) -> int:
        return self.D + 2 * self.D * (self.K - 1) + self.K

    def init_states(self, num_chains: int) -> np.ndarray:
        """Starting values for several chains: chain 0 = `init_params`, the others freshly drawn the same way."""
        D, K = self.D, self.K
        out = np.zeros((num_chains, self.P))
        out[0] = self._flat_state(self.init_params)
        for c in range(1, num_chains):
            out[c] = self._flat_state([np.ones((D, 1)), np.random.normal(0, 1, (D, K - 1)), np.zeros((D, K - 1)),
                                       np.random.normal(0, 1, (K,))])
        return out

    def target_log_prob_fn(self, sigma_d, b_offset, ind_raw, alpha) -> float:
        """Joint log-density at one state (scCODA_model.py:756-760), evaluated by the fp64 device kernel."""
        lp, _ = self.log_prob_and_grad(sigma_d, b_offset, ind_raw, alpha)
        return lp

    def log_prob_and_grad(self, sigma_d, b_offset, ind_raw, alpha, dtype=np.float64):
        """Value and gradient (flat, parameter order of the state vector) of the log-density at one state."""
        prob = self._device_problem(1, dtype)
        theta = self._flat_state([sigma_d, b_offset, ind_raw, alpha])
        lp, g = prob.logp_grad(theta[None, None].astype(dtype))
        return float(lp.item()), g.double().cpu().numpy()[0, 0]

    def get_y_hat(self, states_burnin: List[np.ndarray], num_results: int, num_burnin: int, lazy: bool = False
                  ) -> np.ndarray:
        """Derived quantities per draw — ind, b_raw, beta, concentration, prediction are APPENDED to
        `states_burnin` — and the posterior-mode count matrix (scCODA_model.py:773-840), vectorised.

        lazy=True (what the sample_* methods pass): `concentration` and `prediction` [draw, sample, cell_type] are not
        computed here but appended as `LazyArray`s that the device fills on first use."""
        sigma_d, b_offset, ind_raw, alphas = states_burnin[:4]
        ref = self.reference_cell_type
        s = 50.0 * ind_raw
        e = np.exp(-np.abs(s))
        ind_ = np.where(s >= 0, 1.0 / (1.0 + e), e / (1.0 + e))                   # == exp(s)/(1+exp(s)), overflow-free
        b_raw_ = b_offset * sigma_d                                               # :812
        beta_nb = ind_ * b_raw_                                                   # :814
        beta_ = np.insert(beta_nb, ref, 0.0, axis=-1)                             # :816-820
        lz = self._lazy_predictive(num_burnin, one_chain=alphas.ndim == 2) if lazy else None
        if lz is not None and lz[0].shape[:-2] == alphas.shape[:-1]:
            conc_, pred_ = lz
        else:
            conc_ = np.exp(np.einsum("nd,...dk->...nk", self.x, beta_) + alphas[..., None, :])   # :821-822
            pred_ = self.n_total[:, None] * conc_ / conc_.sum(axis=-1, keepdims=True)          # :824-827 DirMult mean
        states_burnin.extend([ind_, b_raw_, beta_, conc_, pred_])                 # :830-834
        draw_axes = tuple(range(alphas.ndim - 1))
        betas_final = beta_.mean(axis=draw_axes)
        alphas_final = alphas.mean(axis=draw_axes)
        concentration = np.exp(self.x @ betas_final + alphas_final)               # :836
        return concentration / concentration.sum(axis=1, keepdims=True) * self.n_total[:, None]   # :838
