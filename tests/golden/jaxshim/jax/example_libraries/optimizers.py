"""adagrad of jax.example_libraries.optimizers (published update rule)."""
import numpy as np


def adagrad(step_size, momentum=0.9):
    step = step_size if callable(step_size) else (lambda i: step_size)

    def init(x0):
        return x0, np.zeros_like(x0), np.zeros_like(x0)

    def update(i, g, state):
        x, g_sq, m = state
        g_sq = g_sq + np.square(g)
        inv = np.where(g_sq > 0, 1.0 / np.sqrt(np.where(g_sq > 0, g_sq, 1.0)), 0.0)
        m = (1.0 - momentum) * (g * inv) + momentum * m
        return x - step(i) * m, g_sq, m

    def get_params(state):
        return state[0]
    return init, update, get_params
