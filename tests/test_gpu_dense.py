"""The network's Dense layers on the tensor cores (csrc/dense_tc.cu) against the NumPy forward pass."""
import numpy as np
import pytest

from smash_b200.net import Net

pytestmark = pytest.mark.gpu

# TF32 operands (10-bit mantissa), float32 accumulation, against float64 NumPy: absolute error relative to the span of the
# output's bounds (the last layers are sigmoid + MinMaxScale, so the span is the scale of the output)
TOL = 1e-3
SPAN = np.array([1000.0, 1000.0, 100.0, 1000.0])


def _net(nd, neurons, acts, ncv=4, scale=True, seed=3):
    net = Net()
    for k, (n, a) in enumerate(zip(neurons, acts)):
        opt = {"neurons": n, "kernel_initializer": "glorot_uniform"}
        if k == 0:
            opt["input_shape"] = (nd,)
        net.add("dense", opt)
        if a:
            net.add("activation", {"name": a})
    net.add("dense", {"neurons": ncv, "kernel_initializer": "glorot_uniform"})
    net.add("activation", {"name": "sigmoid"})
    if scale:
        net.add("scale", {"bounds": [(1e-6, 1000.0), (1e-6, 1000.0), (-50.0, 50.0), (1e-6, 1000.0)][:ncv]})
    net.compile("adam", {"learning_rate": 0.01}, random_state=seed)
    return net


@pytest.mark.parametrize("acts", [("relu", "relu"), ("tanh", "selu"), ("leaky_relu", "softplus"), ("elu", None)])
def test_mlp_forward_on_tensor_cores(acts):
    # the graph of _ann_optimize.py:143-168 at a size whose leading dimensions need padding (6 -> 150 -> 75 -> 4), 10 007 rows
    rng = np.random.default_rng(7)
    x = rng.uniform(0.0, 1.0, (10007, 6))
    net = _net(6, (150, 75), acts)
    ref = net._predict(x)
    t = {}
    got = net._predict_device(x, timing=t)
    assert got.shape == ref.shape and np.all(np.isfinite(got))
    err = np.abs(got - ref).max(axis=0) / SPAN
    print("mlp", acts, "max err / span per output", err, "device ms", t["ms"])
    assert np.all(err <= TOL), err


def test_mlp_forward_domain_sized():
    # 200 000 rows, 6 -> 512 -> 256 -> 4: the contraction that matters (2 x 200 000 x 512 x 256 flops) runs at tensor-core speed
    rng = np.random.default_rng(9)
    x = rng.uniform(0.0, 1.0, (200000, 6))
    net = _net(6, (512, 256), ("relu", "relu"))
    t = {}
    got = net._predict_device(x, timing=t)
    ref = net._predict(x[:5000])
    err = np.abs(got[:5000] - ref).max(axis=0) / SPAN
    print("domain-sized mlp: device %.3f ms, %.1f TFLOP/s, err %s" % (t["ms"], t["tflops"], err))
    assert np.all(err <= TOL), err
    assert t["tflops"] > 20.0                                              # far above what float32 SIMT code could reach
