"""graphrole_b200: B200-native (sm_100a) hot paths of GraphRole behind GraphRole's Python API.

    from graphrole_b200 import RecursiveFeatureExtractor, RoleExtractor

are drop-ins for the classes exported by graphrole/__init__.py:1-2.  The ReFeX neighbourhood
aggregation and the RolX NMF multiplicative-update loop run as hand-written CUDA kernels in
graphrole_b200/csrc (C-ABI: include/graphrole_b200.h); there is no CPU fallback.
"""
from graphrole_b200.features.extract import RecursiveFeatureExtractor
from graphrole_b200.roles.extract import RoleExtractor
from graphrole_b200.graph.csr import CSRGraph

__all__ = ['RecursiveFeatureExtractor', 'RoleExtractor', 'CSRGraph']
__version__ = '0.1.0'
