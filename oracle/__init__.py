"""CPU oracle for the pySpade hypergeometric DE hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyspade_b200/`` may import this
package.  The only permitted callers are ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py`` (its ``cpu_baseline`` leg and ``--impl reference`` arm), and
there only as the checker / the CPU baseline, never as the thing shipped.

Contents
--------
``ref_loader``      imports the *unmodified* reference from ``/root/reference``
                    (dev container only; the GPU box has no reference tree).
``hypergeom_port``  restatement of ``scipy.stats.hypergeom.logcdf/logsf``
                    (scipy 1.18.1, the un-vendored dependency that holds the
                    p-value arithmetic).
``de_port``         restatement of ``pySpade/utils.py`` hot-path functions
                    (per-gene test, fan-out, CPM normalisation) plus a
                    vectorised all-genes form used for large parity checks.
``hypergeom_oracle.c``  plain-C form of ``hypergeom_port`` (fast bulk checker).

Parity pinning: the reference has no tests and no golden vectors, so the
oracle is pinned against outputs of the reference itself executed in the dev
container (``tests/golden/*.json|npz`` written by ``tests/golden/make_golden.py``)
and against live ``scipy.stats.hypergeom`` (present on the GPU box as well).
"""
