"""
CPU model of the binade scan (cvmatrix_b200/csrc/kernels_scan.cuh) in numpy scalars: the same four passes -
segment sums, prefix + fast / slow decision, proxy chains from +-1.5 * 2^E for both parities, one exact add per
segment - checked bit for bit against the plain sequential chain (what np.sum(A, axis=0) does per column,
cvmatrix/cvmatrix.py:709-737, 1231-1241) on friendly and adversarial inputs.  Runs without a GPU.
"""

import numpy as np
import pytest


def _bits(x):
    return int(np.float64(x).view(np.uint64))


def _from_bits(b):
    return np.uint64(b).view(np.float64)


def scan_sum(x, L=64, s0=0.0, counts=None, p_start=None, tot_start=0.0):
    """s0: the exact chain value before x[0] (pass 4).  p_start / tot_start: what pass 2 uses for it - the same value in the
    plain scan; in the decoupled slab chain (cvmx_slab_scan_local / _prepare) only an APPROXIMATION of it and the sum of
    magnitudes of everything before, because the exact chains of the earlier slabs are not known yet."""
    n = len(x)
    S = (n + L - 1) // L
    segS = [np.sum(x[j * L:(j + 1) * L]) for j in range(S)]            # pass 1 (any order)
    segA = [np.sum(np.abs(x[j * L:(j + 1) * L])) for j in range(S)]
    B0 = [0.0] * S
    ident = [False] * S
    P, tot = np.float64(s0 if p_start is None else p_start), np.float64(tot_start)
    for j in range(S):                                                  # pass 2
        A1 = segA[j] * (1 + 2.0 ** -20)
        margin = 2.0 ** -24 * tot + 2.0 ** -40 * abs(P)
        lo, hi = (P - A1) - margin, (P + A1) + margin
        blo, bhi = _bits(lo), _bits(hi)
        elo, ehi = (blo >> 52) & 0x7FF, (bhi >> 52) & 0x7FF
        if segA[j] == 0.0:
            ident[j] = True
        elif (blo >> 63) == (bhi >> 63) and elo == ehi and 1 <= elo <= 2046:
            B0[j] = _from_bits((blo & 0xFFF0000000000000) | 0x0008000000000000)
        P, tot = P + segS[j], tot + segA[j]
    d0, d1 = [0.0] * S, [np.nan] * S
    for j in range(S):                                                  # pass 3
        if ident[j]:
            c = np.float64(-0.0)
            for v in x[j * L:(j + 1) * L]:
                c = c + v
            d0[j] = d1[j] = c
        elif B0[j] != 0:
            b0 = np.float64(B0[j])
            b1 = _from_bits(_bits(b0) | 1)
            c0, c1 = b0, b1
            for v in x[j * L:(j + 1) * L]:
                c0, c1 = c0 + v, c1 + v
            d0[j], d1[j] = c0 - b0, c1 - b1
    s, nfast = np.float64(s0), 0
    for j in range(S):                                                  # pass 4
        if d1[j] == d1[j]:
            s = s + (d1[j] if (_bits(s) & 1) else d0[j])
            nfast += 1
        else:
            for v in x[j * L:(j + 1) * L]:
                s = s + v
    if counts is not None:
        counts.append((nfast, S))
    return s


def chain_sum(x, s0=0.0):
    s = np.float64(s0)
    for v in x:
        s = s + v
    return s


def _cases():
    rng = np.random.default_rng(1)
    n = 20_000
    yield "uniform w*x", rng.random(n) * rng.random(n), 0.0, True
    yield "uniform squares", rng.random(n) * rng.random(n) * rng.random(n), 0.0, True
    yield "normal (cancelling)", rng.standard_normal(n), 0.0, False
    yield "normal + 3", rng.standard_normal(n) + 3, 0.0, True
    yield "lognormal, 30 binades", np.exp(rng.standard_normal(n) * 8), 0.0, True
    yield "negative", -rng.random(n), 0.0, True
    yield "ties: small multiples of 2^-40 on 1.5", rng.integers(1, 8, n) * 2.0 ** -40, 1.5, True
    yield "ties: half ulps on 1.5", np.full(n, 2.0 ** -53), 1.5, True
    yield "ties: multiples of ulp/2", rng.integers(0, 16, n) * 2.0 ** -53, 1.5, True
    yield "ties: float32-origin values", rng.random(n).astype(np.float32).astype(np.float64), 0.0, True
    yield "integers", rng.integers(0, 1000, n).astype(np.float64), 0.0, True
    yield "huge then small", np.concatenate([[1e300], rng.random(n)]), 0.0, True
    yield "inf inside", np.concatenate([rng.random(5000), [np.inf], rng.random(5000)]), 0.0, True
    yield "nan inside", np.concatenate([rng.random(5000), [np.nan], rng.random(5000)]), 0.0, True
    yield "denormals", rng.random(n) * 5e-321, 0.0, False
    yield "zeros", np.zeros(n), 0.0, True
    yield "negative zeros on -0", np.full(n, -0.0), -0.0, True
    yield "zeros then ones", np.concatenate([np.zeros(1000), np.ones(n)]), 0.0, True
    yield "continued chain", rng.random(n), 12345.678, True
    yield "continued, negative start (crosses zero)", rng.random(n), -5000.25, False
    yield "alternating +-1e10", np.tile([1e10, -1e10 + 1], n // 2) + rng.random(n), 0.0, False
    yield "hover at 2^10", np.concatenate([[1024.0 - 1e-9], rng.standard_normal(n) * 1e-10]), 0.0, False
    yield "creep up from 2^10", np.concatenate([[1024.0], rng.random(n) * 1e-13]), 0.0, False
    yield "creep below 2^10", np.concatenate([[1024.0], -rng.random(n) * 1e-13]), 0.0, False
    for m in (1, 63, 64, 65, 1000):
        yield f"short n={m}", rng.random(m), 0.0, False


@pytest.mark.parametrize("L", [64, 256])
def test_scan_model_is_bit_identical_to_the_sequential_chain(L):
    with np.errstate(all="ignore"):
        for name, x, s0, expect_fast in _cases():
            counts = []
            a, b = chain_sum(x, s0), scan_sum(x, L, s0, counts)
            assert (_bits(a) == _bits(b)) or (a != a and b != b), (name, L, a, b)
            nfast, nseg = counts[0]
            if expect_fast and len(x) >= 20_000:
                assert nfast >= 0.75 * nseg, (name, L, nfast, nseg)   # the scan actually carries these cases


@pytest.mark.parametrize("L", [64, 256])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_decoupled_slab_chain_model(L, world):
    """Row slabs (cvmx_slab_scan_local / _prepare): every slab runs passes 1 - 3 from the SUM OF THE EARLIER SLABS' pass-1
    totals (numpy's pairwise sums here: a different order than the chain, so only approximately the chain's value) and the
    sum of their magnitudes; pass 4 alone continues the exact chain.  The result must still be the sequential chain's, bit
    for bit - a wrong approximation may only cost slow segments."""
    with np.errstate(all="ignore"):
        for name, x, s0, _ in _cases():
            x = np.asarray(x, dtype=np.float64)
            if len(x) < 4 * world:
                continue
            cuts = [len(x) * r // world for r in range(world + 1)]
            slabs = [x[cuts[r]:cuts[r + 1]] for r in range(world)]
            approx = [np.sum(sl) for sl in slabs]                       # pass 1b of every slab, "all-gathered"
            mags = [np.sum(np.abs(sl)) for sl in slabs]
            carry = np.float64(s0)
            for r, sl in enumerate(slabs):
                p_start = np.float64(s0) + np.sum(approx[:r]) if r else np.float64(s0)
                tot_start = np.sum(mags[:r]) + (abs(s0) if r else 0.0)
                carry = scan_sum(sl, L, carry, p_start=p_start, tot_start=tot_start)
            want = chain_sum(x, s0)
            assert (_bits(want) == _bits(carry)) or (want != want and carry != carry), (name, L, world, want, carry)
