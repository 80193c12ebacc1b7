"""Type aliases shared across the package (mirrors flexs/types.py:1-6 of the reference)."""
from typing import List, Union

import numpy as np

#: What ``Landscape.get_fitness`` accepts: a list of strings or a numpy array of strings.
#: The B200 surrogates additionally accept pre-encoded ``uint8[N, L]`` arrays / CUDA tensors.
SEQUENCES_TYPE = Union[List[str], np.ndarray]
