"""
The reference's OWN test-suite (tests/test_cvmatrix.py of sm00thix/cvmatrix v3.2.1, staged unmodified in
oracle/_ref/tests by `make -C oracle ref`) run against the drop-in classes through tests/ref_shim.py, with the
reference's runtime type checking (pyproject.toml:46-53: --typeguard-packages) pointed at cvmatrix_b200.

Expected differences, and the only ones tolerated: `test_dtype` (the sweep includes float16 and float128, which
have no B200 path and are rejected at construction).  JAX-parametrised cases skip (no jax; out of scope).
"""

import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
EXPECTED_FAILURES = {"test_dtype"}


def _run(target, extra=()):
    env = dict(os.environ, CVMX_SHIM_TARGET=target,
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, REF, os.environ.get("PYTHONPATH", "")]))
    cmd = [sys.executable, "-m", "pytest", "tests/test_cvmatrix.py", "-p", "ref_shim", "-o", "addopts=", "-q", "-rfEs",
           "-p", "no:cacheprovider", *extra]
    return subprocess.run(cmd, cwd=REF, env=env, capture_output=True, text=True, timeout=3000)


def _counts(text):
    tail = text.strip().splitlines()[-1] if text.strip() else ""
    return {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|skipped|error|errors)", tail)}


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "tests", "test_cvmatrix.py")), reason="oracle/_ref not staged (make -C oracle ref)")
def test_shim_runs_reference_suite_against_reference_itself():
    """CPU sanity check of the shim: a fast subset of the reference suite passes against the reference package."""
    r = _run("reference", ("-k", "test_errors or test_copy or test_no_response_variables or test_constant_columns or test_invalid_backend"))
    c = _counts(r.stdout)
    assert r.returncode == 0 and c.get("passed", 0) >= 5 and not c.get("failed"), (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.gpu
@pytest.mark.slow
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "tests", "test_cvmatrix.py")), reason="oracle/_ref not staged (make -C oracle ref)")
def test_reference_suite_grades_the_drop_in():
    try:
        import typeguard  # noqa: F401
        tg = ("--typeguard-packages=cvmatrix_b200",)
    except ImportError:
        tg = ()
    r = _run("b200", tg)
    out = r.stdout
    c = _counts(out)
    failed = set(re.findall(r"^FAILED tests/test_cvmatrix\.py::TestClass::(\w+)", out, flags=re.M))
    errors = set(re.findall(r"^ERROR tests/test_cvmatrix\.py::TestClass::(\w+)", out, flags=re.M))
    summary = {"passed": c.get("passed", 0), "failed": c.get("failed", 0), "skipped": c.get("skipped", 0),
               "errors": c.get("error", 0) + c.get("errors", 0), "failed_tests": sorted(failed), "typeguard": bool(tg),
               "expected_failures": sorted(EXPECTED_FAILURES)}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "ref_suite.json"), "w") as f:
            json.dump(summary, f)
        with open(os.path.join(ROOT, "gpurun_out", "ref_suite.txt"), "w") as f:
            f.write(out[-20000:] + "\n---- stderr ----\n" + r.stderr[-5000:])
    except OSError:
        pass
    assert "ref_shim: grading cvmatrix_b200" in r.stderr, r.stderr[-2000:]
    assert not errors and failed <= EXPECTED_FAILURES, (summary, out[-4000:])
    assert summary["passed"] >= 17, summary
