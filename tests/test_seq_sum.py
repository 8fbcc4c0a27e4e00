"""The order-preserving parallel sum behind TreePSRF.mean (TreePSRF.java:264-271): asm_b200/csrc/seq_sum.cuh.

CPU: the header's binade / parity arithmetic, walked the way the kernel walks it, against the plain loop
(tests/c/seq_sum_check.cpp, compiled here with g++).  GPU: asm_seq_sum against numpy's left-to-right accumulate on
inputs that reach every branch of the kernel (ties in both parities, binade crossings, zeros, subnormals, negative and
non-finite values, tile boundaries)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_seq_sum_arithmetic_on_the_host(tmp_path):
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    exe = str(tmp_path / "seq_sum_check")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "c", "seq_sum_check.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.startswith("ok:") and " 0 failures" in r.stdout
    # the point of the exercise: on PSRF-like rows nearly every add takes the parallel path
    tail = r.stdout.strip().splitlines()[-1].split()
    integer_adds, plain_adds = int(tail[2]), int(tail[5])
    assert integer_adds > 100 * plain_adds


def _left_to_right(v):
    # np.add.accumulate on a contiguous float64 vector is the scalar loop s += v[i] (no pairwise blocking)
    v = np.ascontiguousarray(v, dtype=np.float64)
    return np.float64(0.0) if v.size == 0 else np.add.accumulate(v)[-1]


def _cases():
    rng = np.random.default_rng(20241020)
    out = {}
    for n in (0, 1, 2, 63, 64, 65, 2047, 2048, 2049, 4096, 10000, 28285, 50000, 131072):
        num = rng.integers(1_000_000, 1_400_000, n).astype(np.float64)
        den = rng.integers(1_000_000, 1_400_000, n).astype(np.float64)
        out[f"psrf_like_{n}"] = np.sqrt(num / den)
    out["sqrt_int"] = np.sqrt(rng.integers(0, 5_000_000_000, 30000).astype(np.float64))
    # exact ties: the running sum sits at 2^52-ish ulps of 1.0 and elements of 0.5 / 1.5 / 2.5 ulps arrive
    out["half_ulp"] = np.concatenate([np.full(70, 2.0 ** 46), rng.choice([0.5, 1.5, 2.5, 1.0, 0.25, 0.75, 3.0, 0.5000000000000001], 20000)])
    e0 = 3
    small = np.ldexp(1.0, e0 - 40 - rng.integers(0, 16, 12000))
    kinds = rng.integers(0, 6, 12000)
    out["ties_mixed"] = np.where(kinds == 0, np.ldexp(rng.integers(1, 8, 12000).astype(np.float64), e0),
                         np.where(kinds == 1, small, np.where(kinds == 2, 3 * small,
                         np.where(kinds == 3, np.ldexp(1.0, e0) + small, np.where(kinds == 4, 0.0,
                                  np.ldexp(rng.integers(0, 1024, 12000).astype(np.float64), e0 - 10))))))
    wide = np.ldexp(rng.uniform(1, 2, 9000), rng.integers(-60, 60, 9000))
    wide[rng.integers(0, 9000, 500)] = 0.0
    wide[rng.integers(0, 9000, 500)] = -0.0
    wide[rng.integers(0, 9000, 300)] = 5e-324 * rng.integers(1, 1000, 300)
    out["wide"] = wide
    out["tiny"] = np.ldexp(wide, -1000)
    out["huge"] = np.ldexp(wide, 960)
    signed = rng.uniform(0, 3, 7000)
    signed[rng.integers(0, 7000, 140)] *= -1.0
    signed[rng.integers(0, 7000, 140)] *= -1e3
    out["signed"] = signed
    inf = signed.copy()
    inf[4000] = np.inf
    out["inf"] = inf
    nan = signed.copy()
    nan[123] = np.nan
    out["nan"] = nan
    for c in (1.0, 0.5, 3.0, 1.0 / 3, 1e-3, 1024.0):
        out[f"constant_{c}"] = np.full(40000, c)
    return out


CASES = _cases()


@pytest.mark.gpu
@pytest.mark.parametrize("planned", ["1", "0"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_sum_is_the_scalar_loop_bit_for_bit(gpu, monkeypatch, name, planned):
    # planned = "1": the shipped kernel (predicted binades, one pass per tile, the walk for what the stitch rejects);
    # "0": the binade-by-binade walk alone
    monkeypatch.setenv("ASM_SEQSUM_PLANNED", planned)
    v = CASES[name]
    with np.errstate(all="ignore"):
        want = _left_to_right(v)
    with gpu.GpuContext(2, device=0) as ctx:
        got = np.float64(ctx.seq_sum(v))
    if np.isnan(want):
        assert np.isnan(got)
    else:
        assert got.tobytes() == want.tobytes(), f"{name}: {got!r} != {want!r}"


@pytest.mark.gpu
def test_psrf_mean_is_the_sequential_sum_of_its_rows(gpu, multi_chains):
    """The means asm_psrf_eval returns are the left-to-right sums of the rows it returns (TreePSRF.mean), at a row
    count that spans several tiles of the kernel."""
    ch = multi_chains
    with gpu.GpuContext(ch.n_chains, device=0) as ctx:
        ch.upload(ctx)
        rows, mean = ctx.psrf_eval([0] * ch.n_chains, 0, 700, 1, gpu.ASM_RF_FP4, want_rows=True)
    n = rows.shape[1]
    for a in range(ch.n_chains):
        for b in range(ch.n_chains):
            want = _left_to_right(rows[a, :, b]) / np.float64(n)
            assert np.float64(mean[a, b]).tobytes() == np.float64(want).tobytes()
