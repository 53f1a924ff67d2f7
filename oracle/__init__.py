"""CPU oracle for the DAGMA inner-optimisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``midagma_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the thing the GPU path is compared against -- never as a fallback.

Parity status: **pinned against the reference itself**.  The reference
(fbleile/midagma) ships no golden vectors or asserting tests for this path
(SURVEY.md section 4), so the pin is: ``oracle/make_golden.py`` imports the
unmodified reference from ``/root/reference/src`` in the build container,
runs it on seeded synthetic inputs, and commits the traces under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every restatement in
this package against those traces (bit-identical for the numpy linear path),
and, when ``/root/reference`` is present, against the live reference as well.
"""
