"""Helpers shared by the GPU parity tests (product on cuda:0 vs the CPU oracle)."""
import numpy as np
import torch

from oracle import models as omodels, operator as oop, sampler as osmp, sites as osites


def make_rbm(qtx, N, M, dtype, seed=0, scale=0.6):
    """Same random weights in the product model and in the oracle model."""
    net = omodels.RBM.random(N, M, np.float32 if dtype == torch.float32 else np.float64, seed=seed, scale=scale)
    flat = torch.from_numpy(net.params().copy())
    model = qtx.model.RBM_Dense(M, dtype=dtype, params=flat)
    return model, net


def lattice_pair(qtx, kind, L, nparticles=None):
    qtx.sites.Sites._SITES = None
    if kind == "chain":
        return qtx.sites.Chain(L, Nparticles=nparticles), osites.Chain(L, Nparticles=nparticles)
    if kind == "square":
        return qtx.sites.Square(L, Nparticles=nparticles), osites.Square(L, Nparticles=nparticles)
    if kind == "triangular":
        return qtx.sites.Triangular(L, Nparticles=nparticles), osites.Triangular(L, Nparticles=nparticles)
    raise ValueError(kind)


def check(label, err, tol):
    """Assert ``err <= tol`` and append the measured pair to gpurun_out/parity_report.jsonl (the evidence file of
    the parity bar: float64 1e-10, float32 1e-5 -- BASELINE.json north_star)."""
    import json
    import os

    err, tol = float(err), float(tol)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "what": label,
                                "err": err, "tol": tol, "ok": bool(err <= tol)}) + "\n")
    except OSError:
        pass
    assert err <= tol, f"{label}: {err:.3e} > {tol:.1e}"


def sr_step_tolerance(ob, rtol=1e-12):
    """Tolerance of a float64 SR / MinSR step against the oracle: the bar is 1e-10 on a spectrum with a GAP at the
    cut-off rtol * lambda_max (asserted here: no eigenvalue within a factor 1000 of it); both sides then carry the
    eigenvalue error eps * lambda_max on 1 / lambda_k, i.e. eps * lambda_max / lambda_min(kept) relative, which
    exceeds 1e-10 only for kept eigenvalues below 2e-6 lambda_max."""
    ob = np.asarray(ob)
    T = ob @ ob.T if ob.shape[0] < ob.shape[1] else ob.T @ ob
    w = np.linalg.eigvalsh(T)
    lam = np.abs(w).max()
    cut = rtol * lam
    assert not ((np.abs(w) > cut / 1e3) & (np.abs(w) < cut * 1e3)).any(), "spectrum has no gap at the cut-off"
    kept = np.abs(w)[np.abs(w) >= cut * 1e3]
    return max(1e-10, 200 * np.finfo(np.float64).eps * lam / kept.min())


def to_np(t):
    return t.detach().cpu().numpy()


def chains_equal(a, b):
    return (np.asarray(a) == np.asarray(b)).all(axis=1)
