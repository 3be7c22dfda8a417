"""Sampler base class and `run` driver: the host protocol of mocat/src/sample.py.

`run(scenario, sampler, n, random_key, initial_state=None, initial_extra=None, **kwargs) -> cdict`
(sample.py:110-148): startup -> iterate update until termination -> clean_chain -> .time / .summary.
Here the iteration itself lives on the device: a population sampler implements `_run_device`, which enqueues
whole population steps and polls the device control block every `check_every` iterations, so there is no host
round trip inside an iteration.

Option routing (sample.py:24-37, 51-55), kept exactly because user code relies on it: an option whose name is
an attribute of the sampler sets that attribute; at construction every other option becomes a tunable entry of
`sampler.parameters`; at start-up an option additionally overwrites a parameter of the same name (and never
creates one).  `extra.parameters` of a run is completed from `sampler.parameters` wherever the caller left an
entry missing or None (sample.py:64-71).
"""
import copy
from inspect import isclass
from time import time

import numpy as np

from .core import cdict, static_cdict


class Sampler:
    name = None
    max_iter = 10000

    def __init__(self, name=None, **options):
        if name is not None:
            self.name = name
        if not hasattr(self, 'parameters'):
            self.parameters = cdict()
        self._route(options, construct=True)

    def _route(self, options, construct):
        for key, value in options.items():
            known_attribute = hasattr(self, key)
            if known_attribute:
                setattr(self, key, value)
            if construct:
                if not known_attribute:
                    setattr(self.parameters, key, value)
            elif hasattr(self.parameters, key):
                setattr(self.parameters, key, value)

    def __repr__(self):
        return f"mocat.Sampler.{type(self).__name__}"

    def deepcopy(self):
        return copy.deepcopy(self)

    # -- protocol --------------------------------------------------------------------------------------
    def startup(self, scenario, n, initial_state, initial_extra, **options):
        self._route(options, construct=False)
        if not isinstance(self.max_iter, (int, np.integer)):
            raise AttributeError(repr(self) + ' max_iter must be int')
        extra = initial_extra
        extra.iter = getattr(extra, 'iter', 0)
        if not hasattr(extra, 'parameters'):
            extra.parameters = cdict()
        for key, default in vars(self.parameters).items():
            if getattr(extra.parameters, key, None) is None:
                setattr(extra.parameters, key, default)
        return initial_state, extra

    def update(self, scenario, state, extra):
        raise NotImplementedError(f'{self.name} update not initiated')

    def termination_criterion(self, state, extra):
        return extra.iter >= self.max_iter

    def clean_chain(self, scenario, chain_state):
        return chain_state

    def summary(self, scenario, initial_state, initial_extra):          # sample.py:90-107
        report = static_cdict()
        for field, source in (('sampler', self), ('scenario', scenario)):
            label = getattr(source, 'name', None)
            if label is not None:
                setattr(report, field, label)
        report.parameters = self.parameters
        if hasattr(self, 'tuning'):
            report.tuning = self.tuning
        return report

    def _run_device(self, scenario, initial_state, initial_extra):
        """the device loop; provided by the population samplers"""
        raise NotImplementedError(
            f"{self!r}: only population samplers (SMC, SMC-ABC, SVGD) run on the device; serial single-chain "
            "samplers are not a data-parallel path (SURVEY 2a) and there is no CPU fallback")


def run(scenario, sampler, n, random_key, initial_state=None, initial_extra=None, **kwargs):
    sampler = sampler(**kwargs) if isclass(sampler) else sampler
    sampler.n = n
    extra = cdict() if initial_extra is None else initial_extra
    if random_key is not None:
        extra.random_key = random_key
    state, extra = sampler.startup(scenario, n, initial_state, extra, **kwargs)
    report = sampler.summary(scenario, state, extra)
    began = time()
    chain = sampler.clean_chain(scenario, sampler._run_device(scenario, state, extra))   # sample.py:136-140
    chain.time = time() - began
    chain.summary = report
    return chain
