"""Expected shrink of a float32 accumulator that rounds TOWARD ZERO at every accumulate (the behaviour of tcgen05.mma's
TMEM accumulator), simulated in NumPy for several term distributions: the systematic relative error is
-2.2e-8 per accumulate, independent of the distribution.  csrc/resconv_tc.cu (tc_trunc_comp) multiplies the
accumulators by the expected value 1 + 2.2e-8 n.  CPU only; prints one line per case."""
import numpy as np


def trunc32(x):
    y = np.float32(x)
    y = np.where(np.abs(y.astype(np.float64)) > np.abs(x), np.nextafter(y, np.float32(0)), y)
    return y.astype(np.float32)


def run(g, label, per=2.2e-8):
    n, T = g.shape
    exact = g.sum(axis=1)
    acc = np.zeros(n, dtype=np.float32)
    for t in range(T):
        acc = trunc32(acc.astype(np.float64) + g[:, t])
    err = acc.astype(np.float64) - exact
    m = np.mean(err * np.sign(exact)) / np.mean(np.abs(exact))
    err_c = acc.astype(np.float64) * (1 + T * per) - exact
    mc = np.mean(err_c * np.sign(exact)) / np.mean(np.abs(exact))
    print(f"{label:38s} T={T:3d} bias {m:.3e} ({m / T:.2e} per accumulate) rms {np.sqrt(np.mean(err ** 2) / np.mean(exact ** 2)):.2e}"
          f" | corrected: bias {mc:.2e} rms {np.sqrt(np.mean(err_c ** 2) / np.mean(exact ** 2)):.2e}")
    return m / T


def gelu(a):
    return 0.5 * a * (1 + np.tanh(0.79788 * (a + 0.044715 * a ** 3)))


def main():
    rng = np.random.default_rng(1)
    n = 20000
    out = []
    for T in (18, 54, 72):
        a = gelu(rng.standard_normal((n, T, 16))); w = rng.standard_normal((n, T, 16)) / np.sqrt(16 * T)
        out.append(run((a * w).sum(axis=2), "gelu(N) x N weights (random walk)"))
        a = np.abs(rng.standard_normal((n, T, 16))); w = np.abs(rng.standard_normal((n, T, 16)))
        out.append(run((a * w).sum(axis=2), "all positive (coherent)"))
        a = gelu(2 * rng.standard_normal((n, T, 16)) + 1); w = (rng.standard_normal((n, T, 16)) + 0.3) / np.sqrt(16 * T)
        out.append(run((a * w).sum(axis=2), "biased weights (partly coherent)"))
        a = rng.standard_t(3, (n, T, 16)); w = rng.standard_normal((n, T, 16))
        out.append(run((a * w).sum(axis=2), "heavy tailed"))
    T = 16  # per-sample weight gradient: 16 pixel steps x (hi*hi, lo*hi, hi*lo) into ONE accumulator
    a = gelu(rng.standard_normal((n, T, 16))); d = rng.standard_normal((n, T, 16))
    main_ = (a * d).sum(axis=2)
    g = np.zeros((n, 48))
    g[:, 0::3] = main_
    g[:, 1::3] = main_ * 2.0 ** -11 * rng.standard_normal((n, T))
    g[:, 2::3] = main_ * 2.0 ** -11 * rng.standard_normal((n, T))
    out.append(run(g, "wgrad-like 16 x (main, cross, cross)"))
    print(f"per accumulate: min {min(out):.3e} max {max(out):.3e}")


if __name__ == "__main__":
    main()
