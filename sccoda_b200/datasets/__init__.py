"""Bundled data: the Haber et al. (2017) mouse small-intestine cell counts used by the reference's tutorials and
tests (sccoda/datasets/_datasets.py:7-21, haber_counts.csv)."""
import os

import pandas as pd

_HERE = os.path.dirname(os.path.abspath(__file__))


def haber() -> pd.DataFrame:
    """Cell counts of 8 epithelial cell types in 10 samples (4 control, 2 H.poly day 3, 2 H.poly day 10, 2 Salmonella)."""
    return pd.read_csv(os.path.join(_HERE, "haber_counts.csv"))
