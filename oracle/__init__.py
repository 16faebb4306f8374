"""CPU oracle for the MV-KPConv hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package; the product package never does (it fails loudly when its CUDA
library is missing instead of falling back to anything here).

* ``oracle.geom``    - ctypes access to ``liboracle.so`` (plain-C restatement, ``oracle_geom.c``)
                       and to ``_ref/libref.so`` (the unmodified reference C++ core compiled in
                       place from ``/root/reference``, see ``oracle/Makefile``).
* ``oracle.modules`` - torch/numpy fp32 restatements of ``KPConv.forward``, ``FeatureAggregation``,
                       ``group_points``, ``depth2xyz`` + pose and the point-to-pixel 3-NN.
"""
