"""CPU test pinning the ORACLE to the reference's own end-to-end assertions (tests/unit_tests.py:164-265) and to
the captured tutorial outputs: the sequential C restatement of sample_hmc on Haber/Salmonella must
reproduce the expected-sample counts within +-30 cells, call Enterocyte credible, and land in the tutorials'
acceptance-rate band."""
import numpy as np

from conftest import haber_salm
from oracle import c_oracle, np_oracle as o


def test_oracle_reproduces_reference_end_to_end_assertions():
    x, y, names = haber_salm()
    m = o.prepare(x, y, 5)
    rs = np.random.RandomState(1234)
    C = 4
    init = np.zeros((1, C, m.P))
    init[..., 0] = 1.0
    init[..., 1:8] = rs.randn(C, 7)
    init[..., -8:] = rs.randn(C, 8)
    draws, st = c_oracle.sample(m.x[None], m.y[None], m.n_total[None], [m.logp_const], [m.ref], init, o.METHOD_HMC,
                                20000, 5000, seed=5678)
    for c in range(C):
        d = draws[0, c, 5000:]                                             # python-side second burn-in (:205-227)
        der, y_hat = o.derived_draws(d, m)
        alpha_final = np.round(der["alpha"].mean(0), 3)
        inc, nz, thr, final = o.inclusion_and_effects(der["beta"], 0.05)
        a_s, b_s, lfc = o.expected_samples(alpha_final, final.reshape(1, 8), m.y)
        assert not any(np.abs(np.round(y[:4].mean(0)) - np.round(a_s)) > 30)
        assert not any(np.abs(np.round(y[4:].mean(0)) - np.round(b_s[0])) > 30)
        assert final[1] != 0 and abs(final[1] - 1.35) < 0.15                # Enterocyte: tutorials 1.348 / 1.347
        assert inc[1] > 0.95 and inc[5] == 0.0                              # reference column never included
        acc = st["is_accepted"][0, c].mean()
        assert 0.45 < acc < 0.7                                             # tutorials: 49.7 - 60.7 %
    # tutorial intercepts for ref=Goblet differ slightly; with ref=TA (5) the alphas must still be ordered alike
    am = draws[0, :, 5000:, -8:].mean((0, 1))
    assert np.argmax(am) == 6 and np.argmin(am) == 7                        # TA.Early largest, Tuft smallest
