"""Minimal stand-in for gymnasium (absent from the build container) so that the reference's
controller.py, wrappers.py and gym_system can be imported when generating golden vectors."""
from . import spaces
from .envs import registration

_REGISTRY = {}


class Env:
    metadata = {}

    def reset(self, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith('_') or name == 'env':
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)

    def get_wrapper_attr(self, name):
        return getattr(self, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


def make(id, **kwargs):
    import importlib
    if ':' in id:
        mod, id = id.split(':')
        importlib.import_module(mod)
    entry = registration.REGISTRY[id]
    mod, cls = entry.split(':')
    return getattr(importlib.import_module(mod), cls)(**kwargs)
