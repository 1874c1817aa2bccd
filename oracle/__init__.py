"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DrugGEN encoder hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU arm that is timed beside the
GPU number.  ``druggen_b200`` never imports this package.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference
modules from ``/root/reference/src/model/{layers,models,loss}.py`` in the build
container, runs them on seeded inputs and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
(forward, first-order grads, the WGAN-GP double backward and one AdamW step).
"""
