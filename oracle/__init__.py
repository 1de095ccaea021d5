"""oracle/ -- CPU restatement of the bpl-next Dixon-Coles hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under this directory is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / the timed CPU baseline -- never on the GPU product path (``bpl_next_b200`` does not
import it and fails loudly if its CUDA library is missing).

PARITY UNPINNED.  The reference (anguswilliams91/bpl-next) is pure Python over numpyro 0.13.2 /
jax 0.4.24 (``/root/reference/poetry.lock:229-230,421-422``).  Neither jax, jaxlib nor numpyro is
installed or installable in this image (no network, not in /opt/wheelhouse), so the reference
cannot be executed here, and its own tests (``/root/reference/tests``) hold no golden vector or
known-answer value for the log-density, its gradient or the predictive grid (they assert
validity and statistical behaviour only).  The oracle therefore restates the published
numpyro/jax arithmetic line by line (citations in every function) and is pinned only against
(i) the float64 anchors an independent surveyor derivation produced for ``conftest.dummy_data``
(``tests/test_oracle.py``), (ii) finite differences, and (iii) the
distributional assertions of the reference's tests re-run through the product.
"""
