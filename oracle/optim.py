"""Oracle: the optimizers train.py:76-83 can select (CPU, torch).  TEST INFRASTRUCTURE ONLY.

tf.keras.optimizers.Nadam / Adam / SGD are third-party code that is not under /root/reference
(TensorFlow, unpinned: Dockerfile:1; checkpoint era => TF 2.1.x).  Published update rules
restated here; cross-checked in tests against torch.optim.NAdam/Adam/SGD.
PARITY UNPINNED (no TensorFlow in the image).
"""
from __future__ import annotations

from typing import Dict

import torch


class OracleNadam:
    """Keras Nadam (TF 2.1): beta_1 .9, beta_2 .999, eps 1e-7, schedule_decay .004, momentum_cache."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, schedule_decay=0.004):
        self.lr, self.b1, self.b2, self.eps, self.decay = learning_rate, beta_1, beta_2, epsilon, schedule_decay
        self.iter = 0
        self.momentum_cache = 1.0
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}

    def apply_gradients(self, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor]):
        t = self.iter + 1
        mu_t = self.b1 * (1.0 - 0.5 * (0.96 ** (self.decay * t)))
        mu_t1 = self.b1 * (1.0 - 0.5 * (0.96 ** (self.decay * (t + 1))))
        P_t = self.momentum_cache * mu_t
        P_t1 = P_t * mu_t1
        self.momentum_cache = P_t
        for k, p in params.items():
            g = grads[k]
            if k not in self.m:
                self.m[k] = torch.zeros_like(p)
                self.v[k] = torch.zeros_like(p)
            g_hat = g / (1.0 - P_t)
            self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g
            m_hat = self.m[k] / (1.0 - P_t1)
            self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g * g
            v_hat = self.v[k] / (1.0 - self.b2 ** t)
            m_bar = (1.0 - mu_t) * g_hat + mu_t1 * m_hat
            params[k] = p - self.lr * m_bar / (torch.sqrt(v_hat) + self.eps)
        self.iter = t
        return params


class OracleAdam:
    """Keras Adam (TF 2.1): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps), eps 1e-7."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iter = 0
        self.m, self.v = {}, {}

    def apply_gradients(self, params, grads):
        t = self.iter + 1
        lr_t = self.lr * (1.0 - self.b2 ** t) ** 0.5 / (1.0 - self.b1 ** t)
        for k, p in params.items():
            g = grads[k]
            if k not in self.m:
                self.m[k] = torch.zeros_like(p)
                self.v[k] = torch.zeros_like(p)
            self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g * g
            params[k] = p - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)
        self.iter = t
        return params


class OracleSGD:
    def __init__(self, learning_rate=1e-2):
        self.lr = learning_rate
        self.iter = 0

    def apply_gradients(self, params, grads):
        for k in params:
            params[k] = params[k] - self.lr * grads[k]
        self.iter += 1
        return params


def make_optimizer(kind: str, learning_rate: float):
    """train.py:77-83: 'adam' | 'nadam' | anything else -> SGD."""
    if kind == "adam":
        return OracleAdam(learning_rate)
    if kind == "nadam":
        return OracleNadam(learning_rate)
    return OracleSGD(learning_rate)
