"""oracle/ -- CPU restatement of the bpl-next Dixon-Coles hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under this directory is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / the timed CPU baseline -- never on the GPU product path (``bpl_next_b200`` does not
import it and fails loudly if its CUDA library is missing).

PARITY: PINNED TO THE REFERENCE'S SOURCE, NOT TO ITS NUMERICAL LIBRARIES.  The reference
(anguswilliams91/bpl-next) is pure Python over numpyro 0.13.2 / jax 0.4.24
(``/root/reference/poetry.lock:229-230,421-422``).  Neither jax, jaxlib nor numpyro is installed or
installable in this image (no network, not in /opt/wheelhouse), so the reference cannot run as
published, and its own tests (``/root/reference/tests``) hold no golden vector for the log-density,
its gradient or the predictive grid (they assert validity and statistical behaviour only).
What does run is its source: ``oracle/ref_shim.py`` puts stand-ins for the 38 ``jax.numpy`` /
``numpyro`` names the reference uses into ``sys.modules``, imports ``bpl`` unmodified, calls the
real ``fit()`` of each exported class (so the reference's own data preparation runs), intercepts
``MCMC.run`` and traces the captured ``_model`` in float64; the eager predict methods run the same
way.  ``scripts/make_ref_shim_golden.py`` / ``make_ref_shim_grid_golden.py`` wrote the vectors in
``tests/golden/ref_shim_*.npz`` / ``refgrid_*.npz`` from it; the oracle agrees with them to rounding
(1e-16 unweighted, 1e-8...3e-7 where the product's host prep keeps weights / covariates in float32),
``tests/test_golden.py``.  That pins every line of the reference's model, data-preparation and
predict code.  It does NOT pin numpyro's own arithmetic (distribution log-probs, ``biject_to``
transforms and Jacobians, ``LocScaleReparam``, ``plate`` / ``scale`` / ``factor``): the stand-ins
restate it from its documentation, as ``oracle/models.py`` does -- for that layer the checks remain
(i) the float64 anchors of SURVEY.md Appendix E (an independent derivation, ``tests/test_oracle.py``),
(ii) finite differences and (iii) the reference's distributional test assertions re-run through the
product.  The dynamic class is outside the source check (as written it cannot be indexed
consistently).
"""
