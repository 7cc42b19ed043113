"""dune_copasi_b200 -- B200-native CG-P1 diffusion-reaction hot path of DuneCopasi behind a C ABI.

The product is the C++/CUDA library ``libdune_copasi_b200.so`` (include/dune_copasi_b200.h);
this package only holds its sources, the in-tree build and a ctypes binding of the C ABI.
"""
from .capi import (Comm, Config, DcbError, Grid, Model, Operator, Reducer, ReductionError, Solver,  # noqa: F401
                   Stepper, lib)
