"""Shared helpers for the parity tests."""
import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def assert_close(got, ref, tol, what=""):
    """rel-L2 gate with a diagnostic that localises the worst element."""
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, tuple(got.shape), tuple(ref.shape))
    assert torch.isfinite(got).all(), "%s: non-finite values in the CUDA result" % what
    e = rel_l2(got, ref)
    if not e <= tol:
        d = (got - ref).abs()
        idx = torch.nonzero(d == d.max())[0].tolist()
        bad = (d > 0.05 * ref.abs().max()).float().mean().item()
        raise AssertionError("%s: rel-L2 %.3e > %.1e; max|d| %.4g at %s (got %.5g ref %.5g); %.2f%% elems off by >5%% of max"
                             % (what, e, tol, d.max().item(), idx, got[tuple(idx)].item(), ref[tuple(idx)].item(), 100 * bad))
    _log("%-60s rel-L2 %.3e (tol %.0e)" % (what, e, tol))
    return e


def _log(line):
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity.log"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def bf16_round(t):
    return t.to(torch.bfloat16).float()
