"""
pytest plugin (`-p ref_shim`) that lets the REFERENCE's own test-suite (staged unmodified in oracle/_ref/tests by
`make -C oracle ref`) grade a drop-in implementation offline:

* `tests.load_data.load_csv / load_spectra` download the NIR data set at collection time (tests/load_data.py:31-69);
  there is no network here, so they are replaced by seeded synthetic stand-ins of identical shape and columns
  (26 617 rows; 8 binary variety columns, Moisture, Protein, split in {0, 1, 2}; 102 positive pseudo-absorbances);
* with CVMX_SHIM_TARGET=b200 the names `CVMatrix` / `Partitioner` inside `tests.test_cvmatrix` are rebound to
  cvmatrix_b200's classes (the naive oracle `tests.naive_cvmatrix.NaiveCVMatrix` keeps subclassing the reference).
  CVMX_SHIM_TARGET=reference runs the suite against the reference itself (sanity check of the shim).
"""

import os
import sys

import numpy as np

COLUMNS = ["Rye_Midsummer", "Wheat_H1", "Wheat_H3", "Wheat_H4", "Wheat_H5", "Wheat_Halland", "Wheat_Oland", "Wheat_Spelt",
           "Moisture", "Protein", "split"]
N_ROWS, N_CHANNELS = 26617, 102


def _load_csv():
    import pandas as pd

    rng = np.random.default_rng(2024)
    variety = rng.integers(0, 8, size=N_ROWS)
    data = {c: (variety == i).astype(np.float64) for i, c in enumerate(COLUMNS[:8])}
    data["Moisture"] = rng.normal(12.0, 1.5, size=N_ROWS)
    data["Protein"] = rng.normal(11.0, 2.0, size=N_ROWS)
    data["split"] = rng.integers(0, 3, size=N_ROWS).astype(np.float64)
    return pd.DataFrame(data, columns=COLUMNS).astype(np.float64)


def _load_spectra():
    rng = np.random.default_rng(2025)
    return -np.log10(rng.uniform(0.1, 0.9, size=(N_ROWS, N_CHANNELS)))


def pytest_configure(config):
    import tests.load_data as ld   # oracle/_ref/tests (cwd = oracle/_ref)

    ld.load_csv, ld.load_spectra = _load_csv, _load_spectra
    import tests.test_cvmatrix as mod

    target = os.environ.get("CVMX_SHIM_TARGET", "reference")
    if target == "b200":
        import cvmatrix_b200

        mod.CVMatrix, mod.Partitioner = cvmatrix_b200.CVMatrix, cvmatrix_b200.Partitioner
    sys.stderr.write(f"ref_shim: grading {mod.CVMatrix.__module__}.{mod.CVMatrix.__name__}\n")
