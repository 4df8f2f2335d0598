"""bayesn_b200: B200 (sm_100a) implementation of BayeSN's data-parallel hot path behind the
reference's ``SEDmodel`` API.  ``from bayesn_b200 import SEDmodel``."""
__version__ = '0.1.0'


def __getattr__(name):
    if name == 'SEDmodel':
        from .model import SEDmodel
        return SEDmodel
    raise AttributeError(name)
