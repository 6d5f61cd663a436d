"""henbun_b200: B200-native Monte-Carlo ELBO hot path of Henbun behind Henbun's Python surface."""
__version__ = "0.1.0"
