"""
sccoda_b200 — B200-native MCMC engine for scCODA's Dirichlet-multinomial spike-and-slab model.

Drop-in for the sampler path of theislab/scCODA:
    CompositionalAnalysis(...) -> scCODAModel.sample_hmc / sample_hmc_da / sample_nuts -> CAResult
"""
__version__ = "0.1.0"
