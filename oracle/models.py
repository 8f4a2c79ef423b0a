"""Oracle: RBM_Dense and ResConv amplitudes, local updates and log-derivatives.

Test infrastructure.  Amplitudes are (mult, expo) pairs with psi = mult * exp(expo):
LogArray(sign, logabs) for RBM (quantax/nn/activation.py:17-23) and
ScaleArray(significand, exponent) for ResConv (quantax/nn/activation.py:7-14,26-32),
both cast to float64 at the end of the forward (quantax/state/variational.py:266).

Restates
  quantax/model/shallow_nets.py:35-126  (SingleDense / RBM_Dense, init_internal, ref_forward),
  quantax/model/conv_nets.py:26-183     (_ConvBlock, ResConv, final_layer),
  quantax/nn/conv.py:13-68              (ReshapeConv, ConvSymmetrize),
  quantax/symmetry/symmetry.py:386-392  (character weights 1/nsymm),
  quantax/utils/big_array.py:442-462,563-568,616-628 (ScaleArray normalize/mul/sum),
  quantax/state/variational.py:429-491  (log-derivative: d sig / sig, exponent is stop-gradient).
Third-party semantics assumed from published behaviour (parity unpinned): jax.nn.gelu =
tanh approximation; equinox.nn.Conv(padding="SAME", padding_mode="CIRCULAR") = wrap padding
+ cross-correlation, weight [Cout, Cin, kh, kw], bias [Cout, 1, 1]; ravel_pytree order =
(conv1.weight, conv1.bias, conv2.weight, conv2.bias) per block.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# RBM_Dense
# ----------------------------------------------------------------------------
class RBM:
    def __init__(self, W, b):
        self.W = np.ascontiguousarray(W)
        self.b = np.ascontiguousarray(b)
        self.dtype = self.W.dtype
        self.M, self.N = self.W.shape
        self.nparams = self.W.size + self.b.size

    @staticmethod
    def random(N, M, dtype=np.float32, seed=0, scale=None):
        """Random init of the same family as shallow_nets.py:71-74 (LeCun truncated normal times a
        scalar <= 0.99; the jax PRNG stream itself is unpinned)."""
        rng = np.random.default_rng(seed)
        w = rng.standard_normal((M, N)).clip(-2, 2) / 0.87962566103423978 / np.sqrt(N)
        if scale is None:
            scale = 0.5
        return RBM((w * scale).astype(dtype), (0.01 * rng.standard_normal(M)).astype(dtype))

    def params(self):
        return np.concatenate([self.W.ravel(), self.b])

    def init_internal(self, s):
        """shallow_nets.py:81-85: theta = W s + b (model dtype)."""
        return s.astype(self.dtype) @ self.W.T + self.b

    def psi_from_theta(self, theta):
        """shallow_nets.py:104 + activation.py:17-23: prod cosh as (prod sign, sum log|cosh|)."""
        c = np.cosh(theta)
        sign = np.prod(np.sign(c), axis=-1)
        logabs = np.sum(np.log(np.abs(c)), axis=-1, dtype=self.dtype)
        return sign.astype(np.float64), logabs.astype(np.float64)

    def forward(self, s):
        return self.psi_from_theta(self.init_internal(np.asarray(s)))

    def ref_forward(self, s_new, s_old, nflips, theta):
        """shallow_nets.py:87-108.  ``argwhere(size=nflips)`` pads missing indices with 0."""
        ns = s_new.shape[0]
        diff = s_new != s_old
        idx = np.argsort(~diff, axis=1, kind="stable")[:, :nflips]  # first differing sites, ascending
        ar = np.arange(ns)[:, None]
        idx = np.where(diff[ar, idx], idx, 0)  # fill_value 0 of jnp.argwhere(size=nflips)
        sv = s_new[ar, idx].astype(self.dtype)  # [ns, nflips]
        Wc = self.W.T[idx]  # [ns, nflips, M]
        theta_new = theta + 2 * np.einsum("cfm,cf->cm", Wc, sv).astype(self.dtype)
        return self.psi_from_theta(theta_new), theta_new

    def jacobian(self, s):
        """variational.py:447-450 with LogArray: d(logabs) -> O[s, i*N+j] = tanh(theta_i) s_j,
        O[s, M*N+i] = tanh(theta_i); computed in model dtype then cast (variational.py:491)."""
        s = np.asarray(s)
        t = np.tanh(self.init_internal(s))
        Ow = (t[:, :, None] * s.astype(self.dtype)[:, None, :]).reshape(s.shape[0], -1)
        return np.concatenate([Ow, t], axis=1).astype(np.float64)


# ----------------------------------------------------------------------------
# ResConv
# ----------------------------------------------------------------------------
_K0 = 0.7978845608028654  # sqrt(2/pi)
_K1 = 0.044715


def exp_by_scale(zf):
    """nn/activation.py:26-32 for a batch [B, n] of per-sample inputs: (significand [B, n], exponent [B]) with
    exponent = max |x| over the sample."""
    m = np.max(np.abs(zf), axis=1)
    return np.exp(zf - m[:, None]), m


def sinhp1_by_scale(zf):
    """nn/activation.py:7-14 (sinh(x) + 1 in ScaleArray form) for a batch [B, n]."""
    m = np.max(np.abs(zf), axis=1)
    sig = (np.exp(zf - m[:, None]) - np.exp(-zf - m[:, None])) / zf.dtype.type(2) + np.exp(-m)[:, None]
    return sig, m


def gelu(x):
    dt = x.dtype
    u = dt.type(_K0) * (x + dt.type(_K1) * x * x * x)
    return dt.type(0.5) * x * (dt.type(1) + np.tanh(u))


def gelu_grad(x):
    dt = x.dtype
    u = dt.type(_K0) * (x + dt.type(_K1) * x * x * x)
    t = np.tanh(u)
    du = dt.type(_K0) * (dt.type(1) + dt.type(3 * _K1) * x * x)
    return dt.type(0.5) * (dt.type(1) + t) + dt.type(0.5) * x * (dt.type(1) - t * t) * du


def conv_circ(x, W, b):
    """x [B, Cin, H, Wd]; W [Cout, Cin, kh, kw]; cross-correlation with wrap padding."""
    kh, kw = W.shape[2:]
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    out = np.zeros((x.shape[0], W.shape[0]) + x.shape[2:], dtype=x.dtype)
    for dy in range(kh):
        for dx in range(kw):
            xs = np.roll(x, (ph - dy, pw - dx), axis=(2, 3))
            out += np.einsum("oc,bchw->bohw", W[:, :, dy, dx], xs)
    if b is not None:
        out += b.reshape(1, -1, 1, 1)
    return out


def conv_circ_bwd(x, W, dout, need_dx=True):
    kh, kw = W.shape[2:]
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    dW = np.zeros((x.shape[0],) + W.shape, dtype=x.dtype)  # per-sample weight gradient
    dx_ = np.zeros_like(x) if need_dx else None
    for dy in range(kh):
        for dx in range(kw):
            xs = np.roll(x, (ph - dy, pw - dx), axis=(2, 3))
            dW[:, :, :, dy, dx] = np.einsum("bohw,bchw->boc", dout, xs)
            if need_dx:
                dr = np.roll(dout, (dy - ph, dx - pw), axis=(2, 3))
                dx_ += np.einsum("oc,bohw->bchw", W[:, :, dy, dx], dr)
    db = dout.sum(axis=(2, 3))
    return dW, db, dx_


class ResConv:
    """params: list over blocks of dict(w1, b1, w2, b2) (b2 is None in the last block)."""

    def __init__(self, blocks, shape, final="exp", out_complex=False, phase_kernel=None):
        # out_complex: out_dtype = complex128 -> pair_cpl before the final activation (conv_nets.py:165-170,
        # nn/activation.py:75-81); phase_kernel: float32 [N] of a trailing phase layer x * exp(i kernel.s)
        # (nn/sign.py:8-43,62-75, tutorials/triangular.ipynb:120-128)
        self.out_complex = out_complex
        self.phase_kernel = None if phase_kernel is None else np.asarray(phase_kernel, dtype=np.float32)
        self.blocks = blocks
        self.shape = tuple(shape)  # (Lx, Ly); chains use (1, L)
        self.N = int(np.prod(shape))
        self.nblocks = len(blocks)
        self.C = blocks[0]["w1"].shape[0]
        self.dtype = blocks[0]["w1"].dtype
        self.final = final
        self.nparams = sum(v.size for blk in blocks for v in blk.values() if v is not None)

    @staticmethod
    def random(shape, nblocks, channels, kernel_size, dtype=np.float32, seed=0, final="exp", bias_std=0.0,
               out_complex=False, phase_kernel=None):
        """He truncated normal with fan_in = Cin*kh*kw, bias 0 (conv_nets.py:71,
        nn/initializers.py:91-119); bias_std > 0 only to make tests sensitive to the bias path."""
        rng = np.random.default_rng(seed)
        kh = 1 if shape[0] == 1 else kernel_size
        kw = kernel_size
        blocks = []
        for i in range(nblocks):
            blk = {}
            for name, cin, last in (("1", 1 if i == 0 else channels, False), ("2", channels, i == nblocks - 1)):
                fan_in = cin * kh * kw
                w = rng.standard_normal((channels, cin, kh, kw)).clip(-2, 2) / 0.87962566103423978
                blk["w" + name] = (w * np.sqrt(2.0 / fan_in)).astype(dtype)
                blk["b" + name] = None if last else (bias_std * rng.standard_normal(channels)).astype(dtype)
            blocks.append(blk)
        return ResConv(blocks, shape, final, out_complex=out_complex, phase_kernel=phase_kernel)

    def params(self):
        out = []
        for blk in self.blocks:
            for k in ("w1", "b1", "w2", "b2"):
                if blk[k] is not None:
                    out.append(blk[k].ravel())
        return np.concatenate(out)

    # -- forward ------------------------------------------------------------
    def _trunk(self, s, keep=False):
        dt = self.dtype
        x = np.asarray(s).astype(dt).reshape(-1, 1, *self.shape)
        cache = []
        for i, blk in enumerate(self.blocks):
            res = x
            x = x / dt.type(np.sqrt(i + 1))
            if i == 0:
                x = x / dt.type(np.sqrt(2))
                a1 = x
            else:
                a1 = gelu(x)
            h = conv_circ(a1, blk["w1"], blk["b1"])
            a2 = gelu(h)
            y = conv_circ(a2, blk["w2"], blk["b2"])
            if y.shape[1] > res.shape[1]:
                res = np.repeat(res, y.shape[1] // res.shape[1], axis=1)
            if keep:
                cache.append((x, a1, h, a2))
            x = y + res
        z = x / dt.type(np.sqrt(self.nblocks + 1))
        return z, cache

    def _final(self, z):
        """conv_nets.py:167-173 + activation.py:7-14,26-32 + nn/conv.py:61-68."""
        dt = self.dtype
        B = z.shape[0]
        C = self.C
        if self.out_complex:
            z = (z[:, : C // 2] + 1j * z[:, C // 2:]).astype(np.complex128)  # pair_cpl, astype(out_dtype)
            dt = np.dtype(np.complex128)
            C = C // 2
        zf = z.reshape(B, -1)
        if self.final == "exp":
            sig, m = exp_by_scale(zf)
        elif self.final == "sinhp1":
            sig, m = sinhp1_by_scale(zf)
        else:
            raise ValueError(self.final)
        a = sig.reshape(B, C, self.N).mean(axis=1, dtype=dt)  # reshape(-1, nsymm).mean(0)
        if self.out_complex:
            char = 1.0 / self.N  # character.astype(complex128); its ScaleArray exponent is real
            e_char = np.log(char)
            c1 = char * np.exp(0.0 - e_char)
            return np.sum(a * c1, axis=1), m + e_char, sig
        char = dt.type(1.0 / self.N)  # symmetry.py:391, sector 0
        e_char = np.log(char)  # ScaleArray.from_value(character).normalize(): big_array.py:442-451
        c1 = char * np.exp(dt.type(0) - e_char)
        significand = np.sum(a * c1, axis=1, dtype=dt)
        exponent = m + e_char
        return significand, exponent, sig

    def phase(self, s):
        """exp(i * dot(kernel, s)) in float32 -> complex64 (nn/sign.py:31,36)."""
        ph = np.asarray(s).reshape(-1, self.N).astype(np.float32) @ self.phase_kernel
        return np.exp(1j * ph.astype(np.float32)).astype(np.complex64)

    def forward(self, s):
        z, _ = self._trunk(s)
        significand, exponent, _ = self._final(z)
        if self.out_complex:
            if self.phase_kernel is not None:
                significand = significand * self.phase(s)
            return significand.astype(np.complex128), exponent.astype(np.float64)
        return significand.astype(np.float64), exponent.astype(np.float64)

    # -- per-sample log-derivative --------------------------------------------
    def jacobian(self, s):
        dt = self.dtype
        z, cache = self._trunk(s, keep=True)
        B = z.shape[0]
        if self.out_complex:
            # variational.py:461-487: grad of Re(out) and Im(out) w.r.t. the (real, imag) model outputs, then two
            # backward passes; out = significand / sg(significand) + exponent, the phase layer has no parameters
            C2 = self.C // 2
            zc = (z[:, :C2] + 1j * z[:, C2:]).astype(np.complex128).reshape(B, -1)
            _, _, sig = self._final(z)
            if self.final == "exp":
                dsig = sig
            else:
                m = np.max(np.abs(zc), axis=1)[:, None]
                dsig = (np.exp(zc - m) + np.exp(-zc - m)) / 2
            w = (dsig / sig.sum(axis=1, keepdims=True)).reshape(B, C2, *self.shape)
            inv = 1.0 / np.sqrt(self.nblocks + 1)
            seed_re = np.concatenate([w.real, -w.imag], axis=1) * inv
            seed_im = np.concatenate([w.imag, w.real], axis=1) * inv
            return self._backward(cache, seed_re.astype(dt)) + 1j * self._backward(cache, seed_im.astype(dt))
        zf = z.reshape(B, -1)
        _, _, sig = self._final(z)
        if self.final == "exp":
            dsig = sig
        else:
            m = np.max(np.abs(zf), axis=1)[:, None]
            dsig = (np.exp(zf - m) + np.exp(-zf - m)) / dt.type(2)
        w = dsig / sig.sum(axis=1, keepdims=True, dtype=dt)
        dx = (w / dt.type(np.sqrt(self.nblocks + 1))).reshape(z.shape).astype(dt)
        return self._backward(cache, dx)

    def _backward(self, cache, dx):
        dt = self.dtype
        B = dx.shape[0]
        grads = [None] * self.nblocks
        for i in reversed(range(self.nblocks)):
            blk = self.blocks[i]
            x, a1, h, a2 = cache[i]
            dres = dx
            dW2, db2, da2 = conv_circ_bwd(a2, blk["w2"], dx)
            dh = da2 * gelu_grad(h)
            dW1, db1, da1 = conv_circ_bwd(a1, blk["w1"], dh, need_dx=(i > 0))
            grads[i] = (dW1, db1, dW2, db2 if blk["b2"] is not None else None)
            if i > 0:
                dxs = da1 * gelu_grad(x)
                dx = dxs / dt.type(np.sqrt(i + 1)) + dres
        cols = []
        for g in grads:
            for v in g:
                if v is not None:
                    cols.append(v.reshape(B, -1))
        return np.concatenate(cols, axis=1).astype(np.float64)


class RBMConv:
    """RBM_Conv / SingleConv (shallow_nets.py:129-190): psi = prod cosh(Conv(s) + b) with a full-lattice circular
    convolution, eqx.nn.Conv(kernel_size = lattice extent, padding="SAME", padding_mode="CIRCULAR") = wrap-pad by
    ((L-1)//2, L//2) then VALID cross-correlation.  Parameters: K [C, 1, Lx, Ly], b [C]."""

    def __init__(self, K, b, shape):
        self.K, self.b, self.shape = K, b, tuple(shape)
        self.C = K.shape[0]
        self.N = int(np.prod(shape))
        self.dtype = K.dtype
        self.nparams = K.size + b.size

    @staticmethod
    def random(shape, channels, dtype=np.float32, seed=0, scale=0.3):
        rng = np.random.default_rng(seed)
        K = (rng.standard_normal((channels, 1) + tuple(shape)) * scale / np.sqrt(np.prod(shape))).astype(dtype)
        b = (0.1 * rng.standard_normal(channels)).astype(dtype)
        return RBMConv(K, b, shape)

    def params(self):
        return np.concatenate([self.K.ravel(), self.b.ravel()])

    def theta(self, s):
        """[B, C, Lx, Ly] by the direct definition (not through the dense expansion the product uses)."""
        x = np.asarray(s).astype(self.dtype).reshape(-1, 1, *self.shape)
        Lx, Ly = self.shape
        lox, loy = (Lx - 1) // 2, (Ly - 1) // 2
        out = np.zeros((x.shape[0], self.C, Lx, Ly), dtype=self.dtype)
        for dx in range(Lx):
            for dy in range(Ly):
                xs = np.roll(x[:, 0], (lox - dx, loy - dy), axis=(1, 2))  # xs[r] = x[r + d - lo]
                out += self.K[None, :, 0, dx, dy, None, None] * xs[:, None]
        return out + self.b.reshape(1, -1, 1, 1)

    def forward(self, s):
        th = self.theta(s).reshape(len(np.atleast_2d(s)), -1)
        logabs = np.sum(np.log(np.cosh(th)), axis=1, dtype=self.dtype)  # LogArray(sign = 1, logabs)
        return np.ones(th.shape[0]), logabs.astype(np.float64)

    def jacobian(self, s):
        x = np.asarray(s).astype(self.dtype).reshape(-1, 1, *self.shape)
        t = np.tanh(self.theta(s))
        Lx, Ly = self.shape
        lox, loy = (Lx - 1) // 2, (Ly - 1) // 2
        dK = np.zeros((x.shape[0], self.C, Lx, Ly), dtype=self.dtype)
        for dx in range(Lx):
            for dy in range(Ly):
                xs = np.roll(x[:, 0], (lox - dx, loy - dy), axis=(1, 2))
                dK[:, :, dx, dy] = np.einsum("bcxy,bxy->bc", t, xs)
        db = t.sum(axis=(2, 3))
        return np.concatenate([dK.reshape(x.shape[0], -1), db], axis=1).astype(np.float64)


def dense_value(psi):
    mult, expo = psi
    return mult * np.exp(expo)


# ---- parameter-free sign / phase layers (quantax/nn/sign.py:8-75) ---------------------------------------------------
def neel120_kernel(Lx, Ly, triangular_b=False):
    """nn/sign.py:62-75: float32 kernel of the 120-degree Neel phase, flattened in site order."""
    x = 2 * np.arange(Lx)
    y = np.zeros(Ly, dtype=x.dtype) if triangular_b else np.arange(Ly)
    k = (x[:, None] + y[None, :]) % 3
    return (np.pi / 3 * k - np.pi / 6).astype(np.float32).ravel()


def compute_sign(kernel, s, output, neg=False):
    """nn/sign.py:8-43 for a batch of configurations s [ns, N]: phase = dot(kernel, s) in float32."""
    phase = np.asarray(s).astype(np.float32) @ np.asarray(kernel, dtype=np.float32).ravel()
    if output == "sign":
        out = np.sign(np.cos(phase))
    elif output == "phase":
        out = np.exp(1j * phase).astype(np.complex64)
    elif output == "cos":
        out = np.cos(phase)
    else:
        raise ValueError(f"Unknown output type: {output}")
    return -out if neg else out


def neel120_phase(lattice, s):
    """nn/sign.py:62-75."""
    Lx, Ly = lattice.shape[1:]
    return compute_sign(neel120_kernel(Lx, Ly), s, "phase")

