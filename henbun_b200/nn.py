"""Feed-forward networks, mirror of Henbun/nn.py (MatBias :10-32, NeuralNet :34-87).

``MatBias.__call__`` = activation(clip(x w + b)) is ONE CUDA kernel here (GEMM with bias/clip/
activation epilogue); the reference emits matmul, add, clip and the activation as separate TF ops
(nn.py:32,83).  Activations are named by the tf-compat callables in henbun_b200.tf (tf.sigmoid,
tf.nn.relu, tf.tanh) or by string."""
from __future__ import annotations

import numpy as np

from . import ops
from .param import Variable, Parameterized, graph_key
from ._settings import settings


def sigmoid(x, name=None):
    import torch
    return torch.sigmoid(x)


def relu(x, name=None):
    import torch
    return torch.relu(x)


def tanh(x, name=None):
    import torch
    return torch.tanh(x)


sigmoid._hb_act = 'sigmoid'
relu._hb_act = 'relu'
tanh._hb_act = 'tanh'


def _act_name(fn):
    if fn is None:
        return 'none'
    if isinstance(fn, str):
        return fn
    return getattr(fn, '_hb_act', None)


class MatBias(Parameterized):
    def __init__(self, nodes, n_layers=[], mean=0.0, stddev=1.0, variable=Variable,
                 collections=[graph_key.VARIABLES]):
        """w: [*n_layers, in, out], b: [*n_layers, 1, out] (nn.py:23-29)."""
        assert (len(nodes) == 2)
        Parameterized.__init__(self)
        self.w = variable(shape=[nodes[0], nodes[1]], n_layers=n_layers, mean=mean, stddev=stddev,
                          collections=collections)
        self.b = variable(shape=[1, nodes[1]], n_layers=n_layers, mean=mean, stddev=stddev,
                          collections=collections)

    def __call__(self, x, activation=None):
        num = settings.numerics
        return ops.matbias(x, self.w, self.b, act=activation or 'none', clip=bool(num.clip_by_value),
                           lo=num.clip_value_min, hi=num.clip_value_max)


class NeuralNet(Parameterized):
    def __init__(self, nodes, n_layers=[], mean=0.0, stddev=1.0, variable_types=Variable,
                 neuron_types=sigmoid, collections=[graph_key.VARIABLES]):
        Parameterized.__init__(self)
        self.nodes = nodes
        if not isinstance(variable_types, list):
            variable_types = [variable_types for _ in range(len(nodes) - 1)]
        if not isinstance(neuron_types, list):
            self.neuron_types = [neuron_types for _ in range(len(nodes) - 2)]
        else:
            self.neuron_types = neuron_types
        self._matbias_list = []
        for i in range(len(nodes) - 1):
            matbias = MatBias(nodes=[nodes[i], nodes[i + 1]], n_layers=n_layers, mean=mean, stddev=stddev,
                              variable=variable_types[i], collections=collections)
            self._matbias_list.append(matbias)
            setattr(self, 'matbias' + str(i), matbias)

    def __call__(self, x):
        """y_{l+1} = act_l(matbias_l(y_l)), last layer linear (nn.py:73-84); must run in tf_mode."""
        y = x
        for i in range(len(self.nodes) - 2):
            typ = self.neuron_types[i]
            name = _act_name(typ)
            if name is not None:
                y = self._matbias_list[i](y, activation=name)          # fused epilogue
            else:
                y = typ(self._matbias_list[i](y))                       # arbitrary user callable
        return self._matbias_list[-1](y)

    def __getitem__(self, i):
        return self._matbias_list[i]
