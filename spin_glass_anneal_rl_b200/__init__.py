"""B200-native Ising annealing engine: the batched Monte Carlo sweep path of
spin-glass-anneal-rl behind the reference's own Python API."""
__version__ = "0.1.0"
