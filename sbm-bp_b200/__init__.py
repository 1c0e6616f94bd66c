"""sbm-bp on B200: belief propagation for (degree-corrected) SBM inference and EM learning.

The product is ``libsbmbp.so`` (C ABI in ``include/sbmbp.h``: host C++ graph builder + hand-written
sm_100a kernels) and the ``bin/bp`` command line.  This package is the thin Python face of that ABI:

* ``api``        -- ctypes binding; classes named after the reference's (``blockmodel_t``,
                    ``belief_propagation``) with the same method names and argument meaning
* ``generators`` -- synthetic planted SBM / degree-corrected SBM edge lists of the BASELINE shapes
* ``build``      -- compiles the library in-tree with nvcc for sm_100a

There is no CPU fallback: without the built library, or without a B200, engine calls raise.
"""
__version__ = "0.1"
