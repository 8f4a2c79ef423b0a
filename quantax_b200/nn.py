"""Final activations selectable for ResConv (quantax/nn/activation.py:7-32).  The functions are
markers: the arithmetic is fused into the CUDA forward / backward kernels."""


def exp_by_scale(x):  # quantax/nn/activation.py:26-32
    raise RuntimeError("exp_by_scale is evaluated inside the CUDA kernels; pass it as ResConv(final_activation=...)")


def sinhp1_by_scale(x):  # quantax/nn/activation.py:7-14
    raise RuntimeError("sinhp1_by_scale is evaluated inside the CUDA kernels; pass it as ResConv(final_activation=...)")
