"""Tick pass against the row-based passes on a France window: where do they differ?  (GPU diagnostic)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, smash_b200
from smash_b200 import _lib as L
from test_gpu_parity import random_fields
from test_abi import tick_schedule

lib = L.lib()
def run(opts):
    for k, v in opts.items(): lib.smash_b200_set_option(k.encode(), v)
    lib.smash_b200_clear_cache()
    m = cases.france(T=100, sub=(250, 900, 250, 900), ngauge=4)
    random_fields(m, seed=11)
    smash_b200.forward(m.setup, m.mesh, m.input_data, m.parameters, m.parameters.copy(), m.states, m.states.copy(), m.output)
    return m
a = run({"tick_pass": 1, "tick_min_cells": 1000})
b = run({"tick_pass": 0})
qa, qb = np.asarray(a.output.sparse_qsim_domain, np.float64), np.asarray(b.output.sparse_qsim_domain, np.float64)
rel = np.abs(qa - qb) / np.maximum(np.abs(qb), 1e-6)
print("shape", qa.shape, "max rel", rel.max(), "mean signed rel", ((qa - qb) / np.maximum(np.abs(qb), 1e-6)).mean())
info, unit, sigma = tick_schedule(a, 32, 4736, 13)
ntile = info[1]
# sparse index k -> path order j: sparse storage follows path order
percell = rel.max(axis=1) if qa.shape[0] == a.mesh.nac else rel.max(axis=0)
signed = ((qa - qb) / np.maximum(np.abs(qb), 1e-6))
signed = signed.mean(axis=1) if qa.shape[0] == a.mesh.nac else signed.mean(axis=0)
cls = np.where(unit < 0, 2, np.where(unit >= ntile, 3, 1))
fa = None
for c, name in ((1, "tile cells (S/R)"), (3, "reach cells (D)"), (2, "pit cells")):
    sel = cls == c
    if sel.any(): print(name, sel.sum(), "max rel", percell[sel].max(), "mean signed", signed[sel].mean())
top = np.argsort(-percell)[:10]
print("top cells", [(int(j), float(percell[j]), int(cls[j]), int(sigma[j])) for j in top])
print("cost", float(a.output.cost), float(b.output.cost))
# time profile of the worst cell
j = top[0]
row = (qa[j] - qb[j]) / np.maximum(np.abs(qb[j]), 1e-6) if qa.shape[0] == a.mesh.nac else (qa[:, j] - qb[:, j]) / np.maximum(np.abs(qb[:, j]), 1e-6)
print("worst cell rel by step (first 24)", np.round(row[:24], 8))
