"""Type aliases shared across graphrole_b200 (mirrors the vocabulary of graphrole/types.py)."""
from typing import Dict, Hashable, Tuple, Union

import numpy as np
import pandas as pd

Node = Union[int, str]
Edge = Tuple[Node, Node]
VectorLike = Union[np.ndarray, pd.Series]
MatrixLike = Union[pd.DataFrame, np.ndarray]
DataFrameLike = Union[pd.DataFrame, pd.Series]
# what DataFrame.to_dict() returns: {feature: {node: value}}
DataFrameDict = Dict[str, Dict[Hashable, float]]
FactorTuple = Tuple[np.ndarray, np.ndarray]
