"""Descriptor-to-parameter network of ``Model.ann_optimize`` (smash/core/net.py:22-1117, simulation/_ann_optimize.py:20-254).

Host-side restatement: the network is a few hundred weights for a catchment such as Cance (2 -> 18 -> 9 -> 4), its
forward / backward passes are NumPy; what it differentiates through is the hydrological cost, whose gradient with respect
to the distributed parameters comes from ``forward_b`` on the GPU (``_hcost_prime``, net.py:1001-1056).  Layer and
optimiser arithmetic, weight initialisation and the use of the legacy global NumPy generator follow the reference, so a
given ``random_state`` gives the reference's weights."""
from __future__ import annotations

import copy

import numpy as np

from .solver import _mw_forward
from .solver._derived_types import GPARAMETERS_NAME


class _Sigmoid:
    def __call__(self, x):
        return 1 / (1 + np.exp(-x))

    def gradient(self, x):
        return self(x) * (1 - self(x))


class _Softmax:
    def __call__(self, x):
        e = np.exp(x - np.max(x, axis=-1, keepdims=True))
        return e / np.sum(e, axis=-1, keepdims=True)

    def gradient(self, x):
        p = self(x)
        return p * (1 - p)


class _TanH:
    def __call__(self, x):
        return 2 / (1 + np.exp(-2 * x)) - 1

    def gradient(self, x):
        return 1 - np.power(self(x), 2)


class _ReLU:
    def __call__(self, x):
        return np.where(x >= 0, x, 0)

    def gradient(self, x):
        return np.where(x >= 0, 1, 0)


class _LeakyReLU:
    alpha = 0.2

    def __call__(self, x):
        return np.where(x >= 0, x, self.alpha * x)

    def gradient(self, x):
        return np.where(x >= 0, 1, self.alpha)


class _ELU:
    alpha = 0.1

    def __call__(self, x):
        return np.where(x >= 0.0, x, self.alpha * (np.exp(x) - 1))

    def gradient(self, x):
        return np.where(x >= 0.0, 1, self(x) + self.alpha)


class _SELU:
    alpha, scale = 1.6732632423543772848170429916717, 1.0507009873554804934193349852946

    def __call__(self, x):
        return self.scale * np.where(x >= 0.0, x, self.alpha * (np.exp(x) - 1))

    def gradient(self, x):
        return self.scale * np.where(x >= 0.0, 1, self.alpha * np.exp(x))


class _SoftPlus:
    def __call__(self, x):
        return np.log(1 + np.exp(x))

    def gradient(self, x):
        return 1 / (1 + np.exp(-x))


ACTIVATION_FUNC = {"relu": _ReLU, "sigmoid": _Sigmoid, "selu": _SELU, "elu": _ELU, "softmax": _Softmax,
                   "leaky_relu": _LeakyReLU, "tanh": _TanH, "softplus": _SoftPlus}
WB_INITIALIZER = ("uniform", "glorot_uniform", "he_uniform", "normal", "glorot_normal", "he_normal", "zeros")


def _no_unknown(what, opts):
    if opts:
        raise KeyError("Unknown %s options: '%s'" % (what, ", ".join(map(str, opts))))


class Activation:
    """net.py:458-498"""

    def __init__(self, name, **unknown):
        _no_unknown("Activation Layer", unknown)
        self.input_shape = None
        self.activation_name = name
        self._f = ACTIVATION_FUNC[name.lower()]()
        self.trainable = True

    def _set_input_shape(self, shape):
        self.input_shape = shape

    def layer_name(self):
        return "Activation (%s)" % type(self._f).__name__.lstrip("_")

    def _forward_pass(self, x, training=True):
        self.layer_input = x
        return self._f(x)

    def _backward_pass(self, g):
        return g * self._f.gradient(self.layer_input)

    def output_shape(self):
        return self.input_shape

    def n_params(self):
        return 0


class Scale:
    """net.py:501-534, MinMaxScale :831-845"""

    def __init__(self, bounds, **unknown):
        _no_unknown("Scale Layer", unknown)
        self.input_shape = None
        b = np.array(bounds)
        self.lower, self.upper = np.array([x[0] for x in b]), np.array([x[1] for x in b])
        self.trainable = True

    def _set_input_shape(self, shape):
        self.input_shape = shape

    def layer_name(self):
        return "Scale (MinMaxScale)"

    def _forward_pass(self, x, training=True):
        self.layer_input = x
        return self.lower + x * (self.upper - self.lower)

    def _backward_pass(self, g):
        return g * (self.upper - self.lower)

    def output_shape(self):
        return self.input_shape

    def n_params(self):
        return 0


class Dense:
    """net.py:579-688; weights drawn from the legacy global generator in the order weight, bias (_wb_initialization
    :537-576)."""

    def __init__(self, neurons, input_shape=None, kernel_initializer="glorot_uniform", bias_initializer="zeros", **unknown):
        _no_unknown("Dense Layer", unknown)
        self.layer_input, self.input_shape, self.neurons, self.trainable = None, input_shape, neurons, True
        self.weight = self.bias = None
        self.kernel_initializer, self.bias_initializer = kernel_initializer.lower(), bias_initializer.lower()
        for ini in (self.kernel_initializer, self.bias_initializer):
            if ini not in WB_INITIALIZER:
                raise ValueError(f"Unknown initializer: {ini}. Choices {list(WB_INITIALIZER)}")

    def _set_input_shape(self, shape):
        self.input_shape = shape

    def _draw(self, initializer, shape):
        fin, fout = self.input_shape[0], self.neurons
        kind = initializer.split("_")
        if kind[-1] == "uniform":
            limit = np.sqrt(6 / (fin + fout)) if kind[0] == "glorot" else np.sqrt(6 / fin) if kind[0] == "he" else 1 / np.sqrt(fin)
            return np.random.uniform(-limit, limit, shape)
        if kind[-1] == "normal":
            std = np.sqrt(2 / (fin + fout)) if kind[0] == "glorot" else np.sqrt(2 / fin) if kind[0] == "he" else 0.01
            return np.random.normal(0, std, shape)
        return np.zeros(shape)

    def _initialize(self, optimizer):
        self.weight = self._draw(self.kernel_initializer, (self.input_shape[0], self.neurons))
        self.bias = self._draw(self.bias_initializer, (1, self.neurons))
        self._weight_opt, self._bias_opt = copy.copy(optimizer), copy.copy(optimizer)

    def layer_name(self):
        return "Dense"

    def n_params(self):
        return int(np.prod(self.weight.shape) + np.prod(self.bias.shape))

    def _forward_pass(self, x, training=True):
        if training:
            self.layer_input = x
        return x.dot(self.weight) + self.bias

    def _backward_pass(self, g):
        weight = self.weight
        if self.trainable:
            grad_w = self.layer_input.T.dot(g)
            grad_w0 = np.sum(g, axis=0, keepdims=True)
            self.weight = self._weight_opt.update(self.weight, grad_w)
            self.bias = self._bias_opt.update(self.bias, grad_w0)
        return g.dot(weight.T)

    def output_shape(self):
        return (self.neurons,)


class Dropout:
    """net.py:691-727"""

    def __init__(self, drop_rate, **unknown):
        _no_unknown("Dropout Layer", unknown)
        self.drop_rate, self._mask, self.input_shape, self.trainable = drop_rate, None, None, True

    def _set_input_shape(self, shape):
        self.input_shape = shape

    def layer_name(self):
        return "Dropout"

    def _forward_pass(self, x, training=True):
        c = 1 - self.drop_rate
        if training:
            self._mask = np.random.uniform(size=x.shape) > self.drop_rate
            c = self._mask
        return x * c

    def _backward_pass(self, g):
        return g * self._mask

    def output_shape(self):
        return self.input_shape

    def n_params(self):
        return 0


LAYERS = {"dense": Dense, "activation": Activation, "scale": Scale, "dropout": Dropout}


class SGD:
    def __init__(self, learning_rate=0.01, momentum=0, **unknown):
        _no_unknown("SGD optimizer", unknown)
        self.learning_rate, self.momentum, self.w_updt = learning_rate, momentum, None

    def update(self, w, g):
        if self.w_updt is None:
            self.w_updt = np.zeros(np.shape(w))
        self.w_updt = self.momentum * self.w_updt + (1 - self.momentum) * g
        return w - self.learning_rate * self.w_updt


class Adam:
    """net.py:893-938 (the moment estimates are divided by (1 - b), not by (1 - b^t), as in the reference)"""

    def __init__(self, learning_rate=0.001, b1=0.9, b2=0.999, **unknown):
        _no_unknown("Adam optimizer", unknown)
        self.learning_rate, self.eps, self.m, self.v, self.b1, self.b2 = learning_rate, 1e-8, None, None, b1, b2

    def update(self, w, g):
        if self.m is None:
            self.m, self.v = np.zeros(np.shape(g)), np.zeros(np.shape(g))
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * np.power(g, 2)
        m_hat, v_hat = self.m / (1 - self.b1), self.v / (1 - self.b2)
        return w - self.learning_rate * m_hat / (np.sqrt(v_hat) + self.eps)


class Adagrad:
    def __init__(self, learning_rate=0.01, **unknown):
        _no_unknown("Adagrad optimizer", unknown)
        self.learning_rate, self.G, self.eps = learning_rate, None, 1e-8

    def update(self, w, g):
        if self.G is None:
            self.G = np.zeros(np.shape(w))
        self.G += np.power(g, 2)
        return w - self.learning_rate * g / np.sqrt(self.G + self.eps)


class RMSprop:
    def __init__(self, learning_rate=0.001, rho=0.9, **unknown):
        _no_unknown("RMSprop optimizer", unknown)
        self.learning_rate, self.Eg, self.eps, self.rho = learning_rate, None, 1e-8, rho

    def update(self, w, g):
        if self.Eg is None:
            self.Eg = np.zeros(np.shape(g))
        self.Eg = self.rho * self.Eg + (1 - self.rho) * np.power(g, 2)
        return w - self.learning_rate * g / np.sqrt(self.Eg + self.eps)


OPT_FUNC = {"sgd": SGD, "adam": Adam, "adagrad": Adagrad, "rmsprop": RMSprop}


class Net:
    """net.py:22-436"""

    def __init__(self):
        self.layers, self.history = [], {"loss_train": []}
        self._opt = self._optimizer = self._learning_rate = None
        self._compiled = False

    def add(self, layer, options):
        if not isinstance(layer, str):
            raise TypeError("layer argument must be str")
        if layer.lower() not in LAYERS:
            raise ValueError(f"Unknown layer type '{layer}'. Choices: {list(LAYERS)}")
        lay = LAYERS[layer.lower()](**options)
        if not self.layers:
            if "input_shape" not in options:
                raise TypeError("First layer missing required option argument: 'input_shape'")
            if not isinstance(options["input_shape"], tuple):
                raise ValueError(f"input_shape option should be a tuple, not {type(options['input_shape'])}")
        else:
            lay._set_input_shape(self.layers[-1].output_shape())
        self.layers.append(lay)

    def compile(self, optimizer="adam", options=None, random_state=None):
        if not self.layers:
            raise ValueError("The network does not contain layers")
        if not isinstance(optimizer, str):
            raise TypeError("optimizer argument must be str")
        if optimizer.lower() not in OPT_FUNC:
            raise ValueError(f"Unknown optimizer '{optimizer}'. Choices: {list(OPT_FUNC)}")
        if random_state is not None:
            np.random.seed(random_state)
        opt = OPT_FUNC[optimizer.lower()](**(options or {}))
        for layer in self.layers:
            if hasattr(layer, "_initialize"):
                layer._initialize(opt)
        self._compiled, self._optimizer, self._learning_rate = True, optimizer.lower(), opt.learning_rate

    def copy(self):
        return copy.deepcopy(self)

    def n_params(self):
        return sum(layer.n_params() for layer in self.layers)

    def _forward_pass(self, x, training=True):
        for layer in self.layers:
            x = layer._forward_pass(x, training)
        return x

    def _backward_pass(self, loss_grad):
        for layer in reversed(self.layers):
            loss_grad = layer._backward_pass(loss_grad)

    def _predict(self, x):
        return self._forward_pass(x, training=False)

    def _device_chain(self):
        """The leading Dense (+ Activation) layers that smash_b200_mlp_* can run: [(dense, activation layer or None, code)],
        and the index of the first layer after the chain."""
        chain, i = [], 0
        while i < len(self.layers) and isinstance(self.layers[i], Dense):
            code, actl = 0, None
            nxt = self.layers[i + 1] if i + 1 < len(self.layers) else None
            if isinstance(nxt, Activation) and type(nxt._f).__name__ in self._DEVICE_ACT:
                code, actl = self._DEVICE_ACT[type(nxt._f).__name__], nxt
            chain.append((self.layers[i], actl, code))
            i += 2 if code else 1
        return chain, i

    # activation codes of smash_b200_mlp_forward (include/smash_b200.h)
    _DEVICE_ACT = {"_ReLU": 1, "_Sigmoid": 2, "_TanH": 3, "_LeakyReLU": 4, "_ELU": 5, "_SELU": 6, "_SoftPlus": 7}

    def _predict_device(self, x, timing=None):
        """The forward pass of ``_predict`` with the leading Dense (+ Activation) layers on the GPU's tensor cores
        (``smash_b200_mlp_forward``: TF32 tcgen05 GEMMs with bias and activation fused, the whole chain device-resident);
        whatever follows the chain (Scale, Dropout at inference, softmax) is applied on the host to the small result.
        For domain-sized inputs (France: 906 044 rows x 1 554 neurons); TF32 against float64 NumPy: 1e-3 of the output scale.
        ``timing`` (dict) receives ``ms`` and ``tflops`` of the device layers."""
        import ctypes as C

        from . import _lib as L
        chain3, i = self._device_chain()
        chain = [(d, c) for d, _, c in chain3]
        if not chain:
            return self._predict(x)
        x32 = np.ascontiguousarray(x, dtype=np.float32)
        sizes = np.array([x32.shape[1]] + [d.neurons for d, _ in chain], dtype=np.int32)
        ws = [np.ascontiguousarray(d.weight, dtype=np.float32) for d, _ in chain]
        bs = [np.ascontiguousarray(np.ravel(d.bias), dtype=np.float32) for d, _ in chain]
        acts = np.array([a for _, a in chain], dtype=np.int32)
        fpp = C.POINTER(C.c_float) * len(chain)
        y = np.empty((x32.shape[0], int(sizes[-1])), dtype=np.float32)
        ms, fl = C.c_float(0.0), C.c_double(0.0)
        L.check(L.lib().smash_b200_mlp_forward(C.c_int64(x32.shape[0]), len(chain), L._ip(sizes), L._fp(x32), fpp(*[L._fp(w) for w in ws]),
                                               fpp(*[L._fp(b) for b in bs]), L._ip(acts), L._fp(y), C.byref(ms), C.byref(fl)))
        if timing is not None:
            timing["ms"], timing["tflops"] = float(ms.value), float(fl.value) / max(ms.value, 1e-9) * 1e-9
        out = y.astype(np.float64)
        for layer in self.layers[i:]:
            out = layer._forward_pass(out, False)
        return out

    def _fit_d2p(self, x_train, instance, control_vector, mask, parameters_bgd, states_bgd, epochs, early_stopping, verbose,
                 solver=None, device=False):
        """net.py:353-415.  device = True: the Dense / Activation chain runs on the GPU's tensor cores (DeviceChain), forward and
        backward; meant for domain-sized inputs."""
        if not self._compiled:
            raise ValueError("The network has not been compiled yet")
        loss_opt = 0
        dev = DeviceChain(self, x_train) if device else None
        for epo in range(epochs):
            y_pred = dev.forward() if dev else self._forward_pass(x_train)
            loss_grad = _hcost_prime(y_pred, control_vector, mask, instance, parameters_bgd, states_bgd, solver)
            loss = instance.output.cost
            if early_stopping and (loss_opt > loss or epo == 0):
                loss_opt = loss
                for layer in self.layers:
                    if hasattr(layer, "_initialize"):
                        layer._weight, layer._bias = np.copy(layer.weight), np.copy(layer.bias)
            if dev:
                dev.backward(loss_grad)
            else:
                self._backward_pass(loss_grad)
            if verbose:
                print(f"    At epoch    {epo + 1:3}    J ={loss:10.6f}    |proj g| ={np.amax(np.abs(loss_grad)):10.6f}")
            self.history["loss_train"].append(loss)
        if dev:
            dev.close()
        if early_stopping:
            for layer in self.layers:
                if hasattr(layer, "_initialize"):
                    layer.weight, layer.bias = np.copy(layer._weight), np.copy(layer._bias)


class DeviceChain:
    """The network's leading Dense (+ Activation) layers resident on the GPU for a training run (``smash_b200_mlp_create`` /
    ``run_forward`` / ``run_backward``, csrc/dense_tc.cu): the rows go up once, every epoch uploads the weights, runs the
    forward pass on the tensor cores, and -- from the gradient of the loss with respect to the chain's output -- returns
    grad_weight / grad_bias of every Dense layer; the optimiser update stays on the host (Dense._backward_pass, net.py:672-685).
    TF32 operands with float32 accumulation: gradients agree with the float64 NumPy backward pass to about 1e-3 of their
    scale (tests/test_gpu_dense.py)."""

    def __init__(self, net, x):
        import ctypes as C

        from . import _lib as L
        self._C, self._L = C, L
        self.net = net
        self.chain, self.after = net._device_chain()
        if not self.chain:
            raise ValueError("the network does not start with a Dense layer")
        self.x = np.ascontiguousarray(x, dtype=np.float32)
        self.sizes = np.array([self.x.shape[1]] + [d.neurons for d, _, _ in self.chain], dtype=np.int32)
        self.acts = np.array([c for _, _, c in self.chain], dtype=np.int32)
        self.h = C.c_void_p()
        L.check(L.lib().smash_b200_mlp_create(C.c_int64(self.x.shape[0]), len(self.chain), L._ip(self.sizes), L._ip(self.acts), C.byref(self.h)))
        self._first = True
        self.ms_forward = self.ms_backward = 0.0

    def close(self):
        if self.h:
            self._L.lib().smash_b200_mlp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ptrs(self, arrays):
        return (self._C.POINTER(self._C.c_float) * len(arrays))(*[self._L._fp(a) for a in arrays])

    def forward(self):
        C, L = self._C, self._L
        ws = [np.ascontiguousarray(d.weight, dtype=np.float32) for d, _, _ in self.chain]
        bs = [np.ascontiguousarray(np.ravel(d.bias), dtype=np.float32) for d, _, _ in self.chain]
        y = np.empty((self.x.shape[0], int(self.sizes[-1])), dtype=np.float32)
        ms, fl = C.c_float(0.0), C.c_double(0.0)
        L.check(L.lib().smash_b200_mlp_run_forward(self.h, L._fp(self.x) if self._first else None, self._ptrs(ws), self._ptrs(bs), L._fp(y),
                                                   C.byref(ms), C.byref(fl)))
        self._first = False
        self.ms_forward = float(ms.value)
        out = y.astype(np.float64)
        for layer in self.net.layers[self.after:]:
            out = layer._forward_pass(out, True)
        return out

    def backward(self, loss_grad):
        """Net._backward_pass (net.py:301-303): the layers after the chain on the host, the chain on the device, then the
        optimiser updates in the reference's order (last layer first)."""
        C, L = self._C, self._L
        g = loss_grad
        for layer in reversed(self.net.layers[self.after:]):
            g = layer._backward_pass(g)
        gy = np.ascontiguousarray(g, dtype=np.float32)
        gws = [np.zeros((int(self.sizes[l]), int(self.sizes[l + 1])), dtype=np.float32) for l in range(len(self.chain))]
        gbs = [np.zeros(int(self.sizes[l + 1]), dtype=np.float32) for l in range(len(self.chain))]
        ms = C.c_float(0.0)
        L.check(L.lib().smash_b200_mlp_run_backward(self.h, L._fp(gy), self._ptrs(gws), self._ptrs(gbs), C.byref(ms)))
        self.ms_backward = float(ms.value)
        for (dense, _, _), gw, gb in reversed(list(zip(self.chain, gws, gbs))):
            if dense.trainable:
                dense.weight = dense._weight_opt.update(dense.weight, gw.astype(np.float64))
                dense.bias = dense._bias_opt.update(dense.bias, gb.astype(np.float64)[None, :])
        return gws, gbs


def _hcost_prime(y, control_vector, mask, instance, parameters_bgd, states_bgd, solver=None):
    """net.py:1001-1056: predicted fields on the active cells, one ``forward_b``, gradient rows in the same cell order."""
    sv = solver or _mw_forward
    for i, name in enumerate(control_vector):
        getattr(instance.parameters if name in GPARAMETERS_NAME else instance.states, name)[mask] = y[:, i]
    parameters_b, states_b = instance.parameters.copy(), instance.states.copy()
    sv.forward_b(instance.setup, instance.mesh, instance.input_data, instance.parameters, parameters_b, parameters_bgd,
                 instance.parameters.copy(), instance.states, states_b, states_bgd, instance.states.copy(), instance.output,
                 None, 0.0, 1.0)
    return np.transpose([getattr(parameters_b if name in GPARAMETERS_NAME else states_b, name)[mask] for name in control_vector])
