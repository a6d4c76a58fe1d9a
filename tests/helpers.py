"""Shared comparison helpers for the parity tests."""
import numpy as np

SAMPLE_STRIDE = 997  # must match tests/golden/make_golden.py


def rel_err(a, b):
    """max|a-b| / max|b|: the 'relative' error the north_star tolerances are stated in
    (normalised by the tensor's largest magnitude; element-wise ratios are meaningless
    for entries that are ~0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    return float(np.abs(a - b).max() / denom) if b.size else 0.0


def check_summary(golden, prefix, name, x, tol, atol=0.0, robust=False):
    """Compare tensor x with what make_golden.summarize stored under prefix/name.

    robust=True (multi-step optimizer state): Adam turns 1e-9-level gradient noise on elements where the
    gradient and the weight-decay term cancel into lr-sized steps, and a 1e-5 pre-activation difference can
    flip a ReLU for one sample, so a handful of elements legitimately differ by O(lr) between ANY two fp32
    implementations (tools/diag_steps.py quantifies this).  The check is then on the mean absolute error
    and on the fraction of elements off by more than 100*tol, not on the maximum."""
    x = np.asarray(x)
    kf, ks = f"{prefix}/{name}/full", f"{prefix}/{name}/sample"
    if kf in golden:
        ref = golden[kf]
        got = x.reshape(ref.shape)
    else:
        ref = golden[ks]
        got = x.reshape(-1)[::SAMPLE_STRIDE]
    if ref.dtype.kind in "iu":
        assert np.array_equal(got, ref), f"{prefix}/{name}: integer mismatch"
        return 0.0
    n = float(golden[f"{prefix}/{name}/norm"])
    scale = max(float(np.abs(ref).max()), 1e-30)
    diff = np.abs(got.astype(np.float64) - ref)
    if robust:
        err = float(diff.mean() / scale)
        assert err <= tol + atol / scale, f"{prefix}/{name}: mean rel err {err:.3e} > {tol:.1e}"
        frac = float((diff > 100 * tol * scale + atol).mean())
        assert frac <= 0.03, f"{prefix}/{name}: {frac:.3%} of elements off by > {100 * tol:.0e}"
        return err
    err = float(diff.max() / scale)
    assert err <= tol + atol / scale, f"{prefix}/{name}: rel err {err:.3e} > {tol:.1e}"
    gn = float(np.sqrt((x.astype(np.float64) ** 2).sum()))
    if robust:
        tol = 20 * tol
    assert abs(gn - n) <= tol * max(n, 1e-30) * 10 + atol * np.sqrt(x.size) + 1e-30, f"{prefix}/{name}: norm {gn} vs {n}"
    return err


def unpack_mask(bits, shape):
    return np.unpackbits(bits)[: int(np.prod(shape))].reshape(shape)
