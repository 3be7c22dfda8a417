"""Sampler base class and `run` driver: host mirror of mocat/src/sample.py.

`run(scenario, sampler, n, random_key, initial_state=None, initial_extra=None, **kwargs) -> cdict`
(sample.py:110-148): startup -> loop update until termination -> clean_chain -> .time/.summary.
The loop body runs on the device; the host only enqueues kernels and polls the control block every
`check_every` iterations, so there is no host round trip inside an iteration.
"""
import copy
from inspect import isclass
from time import time

import numpy as np

from .core import cdict, static_cdict


class Sampler:
    parameters: cdict
    name: str = None
    max_iter: int = 10000

    def __init__(self, name=None, **kwargs):
        if name is not None:
            self.name = name
        if not hasattr(self, 'parameters'):
            self.parameters = cdict()
        for key, value in kwargs.items():
            if hasattr(self, key):
                setattr(self, key, value)
            else:
                setattr(self.parameters, key, value)

    def __repr__(self):
        return f"mocat.Sampler.{self.__class__.__name__}"

    def deepcopy(self):
        return copy.deepcopy(self)

    def startup(self, scenario, n, initial_state, initial_extra, **kwargs):      # sample.py:46-72
        for key, value in kwargs.items():
            if hasattr(self, key):
                setattr(self, key, value)
            if hasattr(self, 'parameters') and hasattr(self.parameters, key):
                setattr(self.parameters, key, value)
        if not isinstance(self.max_iter, (int, np.integer)):
            raise AttributeError(self.__repr__() + ' max_iter must be int')
        if not hasattr(initial_extra, 'iter'):
            initial_extra.iter = 0
        if not hasattr(initial_extra, 'parameters'):
            initial_extra.parameters = cdict()
        for key, value in self.parameters.__dict__.items():
            if not hasattr(initial_extra.parameters, key) or getattr(initial_extra.parameters, key) is None:
                setattr(initial_extra.parameters, key, value)
        return initial_state, initial_extra

    def update(self, scenario, state, extra):
        raise NotImplementedError(f'{self.name} update not initiated')

    def termination_criterion(self, state, extra):
        return extra.iter >= self.max_iter

    def clean_chain(self, scenario, chain_state):
        return chain_state

    def summary(self, scenario, initial_state, initial_extra):                    # sample.py:90-107
        summ = static_cdict()
        if getattr(self, 'name', None) is not None:
            summ.sampler = self.name
        if getattr(scenario, 'name', None) is not None:
            summ.scenario = scenario.name
        if hasattr(self, 'parameters'):
            summ.parameters = self.parameters
        if hasattr(self, 'tuning'):
            summ.tuning = self.tuning
        return summ

    # device loop, provided by the population samplers
    def _run_device(self, scenario, initial_state, initial_extra):
        raise NotImplementedError(
            f"{self.__repr__()}: only population samplers (SMC, SMC-ABC, SVGD) run on the device; serial "
            "single-chain samplers are not a data-parallel path (SURVEY 2a) and there is no CPU fallback")


def run(scenario, sampler, n, random_key, initial_state=None, initial_extra=None, **kwargs):
    if isclass(sampler):
        sampler = sampler(**kwargs)
    sampler.n = n
    if initial_extra is None:
        initial_extra = cdict()
    if random_key is not None:
        initial_extra.random_key = random_key
    initial_state, initial_extra = sampler.startup(scenario, n, initial_state, initial_extra, **kwargs)
    summary = sampler.summary(scenario, initial_state, initial_extra)
    start = time()
    chain = sampler._run_device(scenario, initial_state, initial_extra)           # sample.py:136-140
    chain = sampler.clean_chain(scenario, chain)
    end = time()
    chain.time = end - start
    chain.summary = summary
    return chain
