"""The network's Dense layers on the tensor cores (csrc/dense_tc.cu) against the NumPy forward pass."""
import numpy as np
import pytest

from smash_b200.net import Net

pytestmark = pytest.mark.gpu

# TF32 operands (10-bit mantissa), float32 accumulation, against float64 NumPy: absolute error relative to the span of the
# output's bounds (the last layers are sigmoid + MinMaxScale, so the span is the scale of the output)
TOL = 1e-3
SPAN = np.array([1000.0, 1000.0, 100.0, 1000.0])
GTOL = 3e-3


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


@pytest.mark.parametrize("acts", [("relu", "tanh"), ("selu", "softplus"), ("leaky_relu", "elu")])
def test_mlp_backward_on_tensor_cores(acts):
    # one training step of the chain on the device against the NumPy backward pass (Dense._backward_pass net.py:672-685):
    # grad_weight = a^T g (contracted over the 20 011 rows), grad_bias = column sums, g pushed back through W^T and the
    # activation derivatives.  TF32 operands: 3e-3 of each gradient's inf-norm (about 0.3 % of the rows sit within TF32 rounding
    # of a kink of relu / leaky_relu / elu / selu, where the device takes the other branch of the derivative).
    from smash_b200.net import Dense, DeviceChain
    rng = np.random.default_rng(5)
    x = rng.uniform(0.0, 1.0, (20011, 6))
    net = _net(6, (150, 75), acts, seed=4)
    ref_net = net.copy()
    # a loss gradient of one sign per field, as a calibration produces (a random-sign gradient would make every sum below a
    # cancellation, and the measure would be dominated by the few rows whose pre-activation lies within TF32 rounding of a kink)
    gy = rng.uniform(0.5, 1.5, (20011, 4)) * np.array([1e-3, -1e-3, 1e-2, 1e-3])
    # NumPy: gradients of every Dense layer, captured before the optimiser update
    y = ref_net._forward_pass(x)
    want, g = [], gy
    for layer in reversed(ref_net.layers):
        if isinstance(layer, Dense):
            want.append((layer.layer_input.T.dot(g), np.sum(g, axis=0)))
            g = g.dot(layer.weight.T)
        else:
            g = layer._backward_pass(g)
    want = want[::-1]
    dev = DeviceChain(net, x)
    yd = dev.forward()
    assert np.all(np.abs(yd - y).max(axis=0) / SPAN <= TOL)
    gws, gbs = dev.backward(gy)
    print("device ms: forward %.3f backward %.3f" % (dev.ms_forward, dev.ms_backward))
    dev.close()
    errs = []
    for l, ((gw, gb), dw, db) in enumerate(zip(want, gws, gbs)):
        ew = np.abs(dw - gw).max() / np.abs(gw).max()
        eb = np.abs(db - gb).max() / np.abs(gb).max()
        print("layer %d: grad_weight err %.2e grad_bias err %.2e (of the inf-norm)" % (l, ew, eb))
        errs.append(max(ew, eb))
    assert max(errs) <= GTOL, errs
    # the optimiser update happened on the host with the device gradients: the weights moved, and by Adam's bounded step
    for a, b in zip([l for l in net.layers if isinstance(l, Dense)], [l for l in ref_net.layers if isinstance(l, Dense)]):
        assert np.abs(a.weight - b.weight).max() > 0 and np.abs(a.weight - b.weight).max() <= 0.0101


def test_ann_optimize_with_the_network_on_the_device():
    # Model.ann_optimize on Cance with device_net = True: the loss must fall like with the NumPy network (same seed, same
    # graph); TF32 rounding makes the two trajectories differ in the last digits only
    import cases
    from smash_b200 import simulation as S
    a, b = cases.cance(T=480), cases.cance(T=480)
    ra = S.ann_optimize(a, epochs=6, random_state=11, return_net=True, device_net=True)
    rb = S.ann_optimize(b, epochs=6, random_state=11, return_net=True)
    la, lb = np.array(ra[1].history["loss_train"]), np.array(rb[1].history["loss_train"])
    print("loss (device net)", la, "loss (numpy net)", lb)
    assert la[-1] < la[0] and np.all(np.abs(la - lb) <= 2e-2 * np.abs(lb))
