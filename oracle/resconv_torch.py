"""ResConv forward of the reference on the HOST CORES through torch's CPU convolution -- TEST / BENCH
INFRASTRUCTURE ONLY (the CPU leg of bench.py; never imported by the product).

Same arithmetic as ``oracle.models.ResConv`` (quantax/model/conv_nets.py:78-92,163-183, nn/conv.py:31-68,
nn/activation.py:7-32): float32, cross-correlation with wrap padding, tanh-form gelu, ``exp(z - max|z|)`` or
``sinh + 1`` in ScaleArray form, channel mean, sum over translations / N.  The NumPy oracle evaluates the
convolutions with ``einsum`` (8 GFLOP/s on 8 cores); the reference itself runs them through XLA's multi-threaded CPU
convolution, so the timed CPU leg uses the multi-threaded convolution torch ships (oneDNN, all host threads) to be a
fair stand-in.  ``tests/test_oracle_cpu.py`` holds this class to the NumPy oracle."""
import numpy as np
import torch
import torch.nn.functional as F


class TorchResConv:
    def __init__(self, net, threads=None):
        """``net``: an ``oracle.models.ResConv`` with real output (float32 or float64 parameters)."""
        if net.out_complex:
            raise ValueError("TorchResConv covers real-output networks (configs C and E)")
        if threads:
            torch.set_num_threads(int(threads))
        self.net = net
        self.tdt = torch.float32 if net.dtype == np.float32 else torch.float64
        self.npdt = np.dtype(net.dtype).type
        self.shape, self.N, self.C, self.nblocks, self.final = net.shape, net.N, net.C, net.nblocks, net.final
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
        self.blocks = [{k: t(v) for k, v in blk.items()} for blk in net.blocks]

    @staticmethod
    def _conv(x, w, b):
        ph, pw = (w.shape[2] - 1) // 2, (w.shape[3] - 1) // 2
        return F.conv2d(F.pad(x, (pw, pw, ph, ph), mode="circular"), w, b)

    def forward(self, s):
        with torch.no_grad():
            x = torch.from_numpy(np.ascontiguousarray(s)).to(self.tdt).reshape(-1, 1, *self.shape)
            for i, blk in enumerate(self.blocks):
                res = x
                x = x / self.npdt(np.sqrt(i + 1))
                a1 = x / self.npdt(np.sqrt(2)) if i == 0 else F.gelu(x, approximate="tanh")
                h = self._conv(a1, blk["w1"], blk["b1"])
                y = self._conv(F.gelu(h, approximate="tanh"), blk["w2"], blk["b2"])
                if y.shape[1] > res.shape[1]:
                    res = res.repeat_interleave(y.shape[1] // res.shape[1], dim=1)
                x = y + res
            z = (x / self.npdt(np.sqrt(self.nblocks + 1))).reshape(x.shape[0], -1)
            m = z.abs().amax(dim=1)
            if self.final == "exp":
                sig = torch.exp(z - m[:, None])
            else:
                sig = (torch.exp(z - m[:, None]) - torch.exp(-z - m[:, None])) / 2 + torch.exp(-m)[:, None]
            a = sig.reshape(-1, self.C, self.N).mean(dim=1)
            char = self.npdt(1.0 / self.N)
            e_char = np.log(char)
            significand = (a * (char * np.exp(self.npdt(0) - e_char))).sum(dim=1)
            exponent = m + e_char
        return significand.double().numpy(), exponent.double().numpy()
