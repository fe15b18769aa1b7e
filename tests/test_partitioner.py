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
