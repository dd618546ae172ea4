"""TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32 / numpy) of the reference's FRESCO hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker or
as the timed CPU baseline -- never as part of the product path.

Parity pinning: every function here is checked against golden vectors produced
by importing the *real* reference (``/root/reference``) in the build container
(``tests/golden/make_golden.py``; fixtures committed under ``tests/golden``).
"""
