"""Test-infrastructure oracles (CPU restatements of the reference hot path). Never imported by sccoda_b200."""
