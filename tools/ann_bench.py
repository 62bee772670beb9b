"""France-scale forward pass of the ANN mapping on the tensor cores (smash_b200_mlp_forward): device time and TFLOP/s.
usage: python tools/ann_bench.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from smash_b200.net import Net
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 906044
n1 = int(round(np.sqrt(rows * 6) * 2 / 3))
net = Net()
net.add("dense", {"input_shape": (6,), "neurons": n1, "kernel_initializer": "glorot_uniform"})
net.add("activation", {"name": "relu"})
net.add("dense", {"neurons": round(n1 / 2), "kernel_initializer": "glorot_uniform"})
net.add("activation", {"name": "relu"})
net.add("dense", {"neurons": 4, "kernel_initializer": "glorot_uniform"})
net.add("activation", {"name": "sigmoid"})
net.compile("adam", {"learning_rate": 0.003}, random_state=11)
x = np.random.default_rng(3).uniform(0.0, 1.0, (rows, 6)).astype(np.float32)
for _ in range(3):
    t = {}
    y = net._predict_device(x, timing=t)
    print("rows %d graph 6-%d-%d-4: %.3f ms, %.1f TFLOP/s" % (rows, n1, round(n1 / 2), t["ms"], t["tflops"]), flush=True)
ref = net._predict(x[:2000].astype(np.float64))
print("max abs err vs numpy f64 (first 2000 rows):", float(np.abs(y[:2000] - ref).max()))

# one training step with the chain resident on the device: forward + backward (grad_weight, grad_bias of every Dense layer)
from smash_b200.net import DeviceChain
dev = DeviceChain(net, x)
gy = np.random.default_rng(4).uniform(0.5, 1.5, (rows, 4)).astype(np.float32) * 1e-6
for _ in range(3):
    dev.forward()
    dev.backward(gy)
    fl = 2.0 * rows * (6 * n1 + n1 * round(n1 / 2) + round(n1 / 2) * 4)
    print("training step: forward %.3f ms, backward %.3f ms (%.1f TFLOP/s over 2 contractions per layer)" %
          (dev.ms_forward, dev.ms_backward, 2 * fl / dev.ms_backward * 1e-9), flush=True)
dev.close()
