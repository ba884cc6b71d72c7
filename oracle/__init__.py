"""CPU oracle for the WSI multimodal-MIL attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the CPU comparator, never as the thing shipped or measured.

It is a from-scratch functional restatement (plain torch ops on CPU, fp32 or
fp64) of the reference algorithms named in SURVEY.md section 8(a):

* ``oracle.deform1d``  - DeformCrossAttention1D (models/DeformableAttention1D.py:36-240)
* ``oracle.nystrom``   - NystromAttention + moore_penrose_iter_pinv
                         (models/NystromAttention.py:20-157; the pip package
                         ``nystrom_attention`` (lucidrains, version unpinned by the
                         reference) is absent, the vendored copy is the spec)
* ``oracle.towers``    - DeformCrossTransMIL, TransMIL, DeformPathomicNet callers
                         (models/DeformCrossTransMIL.py:28-161, models/mil.py:171-259,
                         models/model.py:173-218,471-568)

Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
imported in the build container from /root/reference by
``oracle/make_goldens.py``; the vectors live in ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks the restatement against them.
"""
