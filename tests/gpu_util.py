"""Helpers for the -m gpu parity tests."""
import torch

from pytorch_stable_diffusion_b200 import _ext


def setup_exact_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel_err(got, ref):
    """max|got - ref| / max|ref| — the per-step metric of BASELINE.json's north_star."""
    got = got.float()
    ref = ref.float()
    denom = ref.abs().max().clamp_min(1e-20)
    return ((got - ref).abs().max() / denom).item()


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs()
    denom = ref.abs().max().clamp_min(1e-20)
    e = (err.max() / denom).item()
    bad = err > tol * denom
    nbad = int(bad.sum().item())
    msg = f"[{name}] rel_err={e:.3e} (tol {tol:.1e}) bad={nbad}/{got.numel()} nan={int(torch.isnan(got).sum())}"
    if nbad:
        idx = bad.nonzero()
        msg += f" first_bad={idx[0].tolist()} last_bad={idx[-1].tolist()}"
        if got.dim() == 2:
            rows = bad.any(dim=1).nonzero().flatten()
            cols = bad.any(dim=0).nonzero().flatten()
            msg += f" bad_rows={rows.numel()} [{rows[:6].tolist()}..] bad_cols={cols.numel()} [{cols[:6].tolist()}..]"
        i0 = tuple(idx[0].tolist())
        msg += f" got={got[i0].item():.5f} ref={ref[i0].item():.5f}"
    print(msg, flush=True)
    fault = _ext.read_fault()
    assert fault == 0, f"{name}: device watchdog fault 0x{fault:x} (site {fault >> 8})"
    assert e <= tol, msg
    return e
