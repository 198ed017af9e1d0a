"""CPU oracle for the GolemFlavor hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``golemflavor_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker (or as the CPU
arm that is timed *beside* the CUDA path), never as the thing shipped.

Parity pinning: ``oracle.golem_oracle`` is checked in ``tests/test_oracle_golden.py``
against (i) every docstring known-answer vector of the reference
(``golemflavor/fr.py``) and (ii) ``tests/golden/*.npz`` fixtures produced by
importing the UNMODIFIED reference in the build container
(``tests/golden/make_golden.py``).
"""
