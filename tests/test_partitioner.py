"""CPU tests of the host-side Partitioner (CSR builder) against the golden fixtures and the oracle."""

import numpy as np
import pytest

import golden_io
from cvmatrix_oracle import OraclePartitioner
from cvmatrix_b200.partitioner import Partitioner


def _same(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].dtype == np.int64 and np.array_equal(a[k], b[k])


def test_golden():
    for pc in golden_io.manifest()["partitioner"]:
        folds = eval(pc["folds_repr"])  # noqa: S307
        d = Partitioner(folds).folds_dict
        assert [repr(k) for k in d] == pc["keys_repr"]
        for v, g in zip(d.values(), pc["indices"]):
            assert v.dtype == np.int64 and v.tolist() == g


@pytest.mark.parametrize("n,p", [(1, 1), (10, 3), (1000, 7), (1000, 1000), (100003, 5), (50000, 50000)])
def test_numeric_fast_path_equals_dict_semantics(n, p):
    rng = np.random.default_rng(n + p)
    for labels in (np.arange(n) % p, rng.integers(-5, p, size=n), rng.permutation(n) % p, (np.arange(n) % p).astype(np.float64),
                   (np.arange(n) % 2).astype(bool), (rng.integers(0, p, size=n)).astype(np.int8 if p < 100 else np.int32)):
        fast = Partitioner(labels)
        _same(fast.folds_dict, OraclePartitioner(labels).folds_dict)
        _same(fast.folds_dict, Partitioner(labels.tolist()).folds_dict)
        off, idx = fast.csr()
        assert off[0] == 0 and off[-1] == n and idx.dtype == np.int64 and idx.flags.c_contiguous
        assert np.array_equal(np.sort(idx), np.arange(n))
        for pos, k in enumerate(fast.folds_dict):
            assert fast.fold_position(k) == pos
            assert np.shares_memory(fast.folds_dict[k], idx) or fast.folds_dict[k].size == 0
        # key types match what iterating the array yields (numpy scalars)
        assert all(type(a) is type(b) for a, b in zip(fast.folds_dict, OraclePartitioner(labels).folds_dict))


def test_special_labels():
    # NaN labels are distinct keys under dict semantics -> hashable path
    labels = np.array([0.0, np.nan, 1.0, np.nan, -0.0])
    a, b = Partitioner(labels).folds_dict, OraclePartitioner(labels).folds_dict
    assert len(a) == len(b) == 4
    for (ka, va), (kb, vb) in zip(a.items(), b.items()):
        assert (ka == kb or (ka != ka and kb != kb)) and np.array_equal(va, vb)
    mixed = [0, "one", 2, 2, "one", 1.0, True, (1, 2), None, (1, 2)]
    _same(Partitioner(mixed).folds_dict, OraclePartitioner(mixed).folds_dict)
    _same(Partitioner(iter(range(5))).folds_dict, OraclePartitioner(range(5)).folds_dict)
    assert Partitioner([]).folds_dict == {} and Partitioner(np.zeros(0, int)).n_folds == 0
    with pytest.raises(ValueError, match="Fold nope not found."):
        Partitioner([1, 2]).get_validation_indices("nope")
    with pytest.raises(TypeError):
        Partitioner(np.zeros((3, 2)))  # rows are unhashable, as in the reference


def test_native_partition_thread_counts_agree(monkeypatch):
    """The threaded counting passes (cvmx_partition_labels, long arrays with few labels) give the serial result for every
    thread count - offsets, indices, key order - including labels that first appear late and negative labels."""
    import numpy as np

    from cvmatrix_b200 import Partitioner

    rng = np.random.default_rng(3)
    N = 250_000
    cases = [np.arange(N) % 5, rng.integers(-3, 40, size=N), np.where(np.arange(N) < N - 7, np.arange(N) % 3, 1000 + np.arange(N) % 2)]
    for labels in cases:
        monkeypatch.setenv("CVMX_PARTITION_THREADS", "1")
        ref = Partitioner(labels)
        for nt in ("2", "3", "8", "13"):
            monkeypatch.setenv("CVMX_PARTITION_THREADS", nt)
            p = Partitioner(labels)
            assert [int(k) for k in p.folds_dict] == [int(k) for k in ref.folds_dict]
            assert np.array_equal(p.offsets, ref.offsets) and np.array_equal(p.indices, ref.indices)
        # against the definition
        first_seen = list(dict.fromkeys(labels.tolist()))
        assert [int(k) for k in ref.folds_dict] == first_seen
        for k in first_seen[:3]:
            assert np.array_equal(ref.get_validation_indices(k), np.flatnonzero(labels == k))
