"""models/functions/funcs.py — only the helper the hot path's weight init needs."""
import math


def bias_init_with_prob(prior_prob):
    """Initial bias so that sigmoid(bias) == prior_prob (models/functions/funcs.py:329-332)."""
    return float(-math.log((1 - prior_prob) / prior_prob))
