"""Optimizer hyper-parameter holder standing in for ``tf.train.AdamOptimizer`` (Henbun/model.py:206).
The update itself is the fused CUDA kernel hb_adam_tf1 (TF-1 rule: lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
epsilon outside the square root)."""


class AdamOptimizer(object):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, use_locking=False, name='Adam'):
        self.learning_rate = float(learning_rate)
        self.beta1 = float(beta1)
        self.beta2 = float(beta2)
        self.epsilon = float(epsilon)
        self.name = name
