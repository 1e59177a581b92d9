"""Adapter registry: same dispatch rule as graphrole/graph/interface/__init__.py:12-53 --
the top-level package name of the graph object's class picks the adapter."""
from typing import Callable, List, Optional

from graphrole_b200.graph.interface.base import BaseGraphInterface
from graphrole_b200.graph.interface.csr import CSRInterface
from graphrole_b200.graph.interface.networkx import NetworkxInterface


# The reference also registers an igraph adapter (graphrole/graph/interface/igraph.py); igraph is
# an optional dependency that is not installed here, SURVEY.md section 2 marks it out of scope, and
# an adapter nobody can run is not shipped.  Any graph library plugs in the reference's way:
# subclass BaseGraphInterface (get_nodes / get_neighbors are enough for the CSR ingest) and
# register it under the graph object's top-level module name.
INTERFACES = {
    'networkx': NetworkxInterface,
    # graphrole_b200.graph.csr.CSRGraph: arrays already in CSR form (and possibly in HBM)
    'graphrole_b200': CSRInterface,
}


def get_supported_graph_libraries() -> List[str]:
    return list(INTERFACES.keys())


def get_interface(G) -> Optional[Callable[..., BaseGraphInterface]]:
    """Adapter factory for graph object G, or None when its library is not supported."""
    module = getattr(G, '__module__', None)
    if not isinstance(module, str) or not module:
        return None
    return INTERFACES.get(module.split('.')[0])
