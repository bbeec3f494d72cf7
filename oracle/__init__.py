"""CPU oracle for the episodic prototypical-network hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package imports this
directory; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

Every function is a CPU restatement (torch-CPU / numpy, fp32 unless noted) of
one reference function and cites the reference ``file:line`` it follows
(paths relative to the reference checkout).  The restatement is pinned against
outputs of the *real* reference, executed in the build container by
``tests/golden/make_golden.py`` and committed as fixtures under
``tests/golden/`` (``tests/test_oracle_golden.py`` re-checks on every run).

Parity status
-------------
* prototypes / proto loss / CPL / SpecAugment / majority vote / view fusion /
  projection / encoders: **pinned** to the reference's own code via fixtures.
* angular loss (``oracle.angular``): **parity unpinned** - the arithmetic lives
  in the third-party ``pytorch_metric_learning`` package, which the reference
  imports (loops/loss.py:5-6) but neither vendors nor pins, and which is not
  installed here.  ``oracle/angular.py`` restates its published algorithm.
"""
