"""ORACLE — test infrastructure only (never imported by the product package).

CPU restatement of the reference hot path used as the parity checker:

* ``pointnet2_ops_ref.c``  – plain C restatement of the nine ``pointnet2_ops._ext`` ops
  (third-party, un-vendored dependency of the reference; see the header of that file).
* ``ops.py``               – ctypes/torch-CPU wrappers around the C library.
* ``ext_stub.py``          – an object with the ``_ext`` call signature, so the reference's
  own ``pointnet2/utils/*.py`` can be executed on CPU when generating golden vectors.
* ``modules.py``           – functional torch-CPU restatement of the reference composition
  (QueryAndGroup, SA / FP modules, backbone, xcorr, RPN, BAT / P2B forward, losses).

PARITY UNPINNED for the nine ops (no upstream vectors exist); the composition is pinned
against the reference's own Python by ``tests/golden/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
