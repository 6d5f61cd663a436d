from . import kernels
from .gp import GP, SparseGP
kern = kernels          # BASELINE.json's wording: gp.kern.RBF / gp.kern.Stationary
